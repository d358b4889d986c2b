// qmf_anal_kernel.cu — fixed-point complex ("HQ") 32-band SBR QMF analysis for sm_100a (B200).
//
// One warp owns one unit (one frame of one core channel: 1024 PCM16 samples -> 32 time slots x 32 complex bands).
// Replaces, bit-exactly, the reference stage
//   ixheaacd_cplx_anal_qmffilt            decoder/generic/ixheaacd_qmf_dec_generic.c:590-741   (HQ, non-ELD path)
// and its leaves
//   ixheaacd_sbr_qmfanal32_winadd         generic:528-588    (5-tap polyphase window, 64 outputs per slot)
//   ixheaacd_fwd_modulation               generic:468-526    (>>4, fold, cos_sin_mod M = 16, t_cos rotation)
//   ixheaacd_cos_sin_mod / radix4bfly / postradixcompute4     generic:259-466, :1736-1829, :1831-1932
//
// The 320-sample WORD16 ring (anal_filter_states) is emulated literally in shared memory, slot by slot, so every
// (position, coefficient-phase) state the reference function accepts is reproduced.  Four slots are modulated together
// so that the 4 x 8 radix-4 butterflies of a 16-point FFT stage fill the warp.
//
// No add in the window or in the modulation can saturate: sum|coef| over the 5 taps <= 32757 so |window output| < 2^30,
// after >>4 and the fold |input| < 2^27, and the modulation grows magnitudes by < 11.4x (bound computed from the tables
// at install time, `anal_exact` below) — those stages use wrapping adds, which is bit-identical. The final t_cos
// rotation can exceed 31 bits and keeps the reference's saturating adds.
// Algorithmic HBM bytes per unit: 2048 (PCM16 in) + 640 + 640 (ring in/out) + 8192 (32 x 32 complex out) = 11520.
#include <cstdint>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"
#include "sbr_glue_units.cuh"

namespace xb {

constexpr int kAnaWarps = 8;
constexpr int kAHalf = 20;            // int2 per 16-point half (16 used + 4 pad)
constexpr int kASlot = 2 * kAHalf;

struct AnaWarpSmem {
  int16_t ring[320];
  int2 T[4 * kASlot];   // FFT workspace for four slots
  int2 F[4 * kASlot];   // FFT output (natural order) for four slots
};

struct AnaBlockSmem {
  int16_t qmf_c[1280];  // prototype, reference order
  int2 pre_tw[16];      // (wim<<16, wre<<16)  sbr_sin_cos_twiddle_l32
  int2 alt_tw[8];       // (wim<<16, wre<<16)  sbr_alt_sin_twiddle_l32
  int2 w[12];           // radix-4: position i -> (si,co) x 3, each << 16   (w_16)
  int2 tcos[32];        // (cosh, sinh) sign-extended                         (sbr_t_cos_sin_l32)
  i32 digrev[2];        // dig_rev_table4_16 >> 3 : complex offset of each output quad
  AnaWarpSmem w_[kAnaWarps];
};

XB_DEV i32 mul32x16_shl(i32 a, i32 c16) { return lsl(__mulhi(a, (i32)((u32)c16 << 16)), 1); }  // ops40.h:23

// FRONT: the unit runs inside sbr_front_hq_kernel — usb comes from the caller (sbr_pre_unit of the same warp) and the
// return value is this lane's OR of ixheaac_abs32_nrm over everything it stored for bands < usb (the headroom scan of
// rows 6..37, ixheaacd_expsubbandsamples at sbr_dec.c:1050, collected on the fly).
template <bool SAT, bool FRONT, bool W32>
__device__ __forceinline__ i32 anal_unit(const QmfAnalArgs &p, AnaBlockSmem &sm, AnaWarpSmem &ws, long long u,
                                         int lane, int usb_in) {
  auto ADD = [](i32 a, i32 b) { return SAT ? add_sat(a, b) : wadd(a, b); };
  auto SUB = [](i32 a, i32 b) { return SAT ? sub_sat(a, b) : wsub(a, b); };
  auto NEG = [](i32 a) { return SAT ? neg_sat(a) : wneg(a); };
  const unsigned full = 0xffffffffu;
  const int16_t *pcm = p.pcm + (p.pcm_unit_stride ? u * p.pcm_unit_stride
                                                     : ((p.ch_fac == 1) ? u * 1024 : (u / p.ch_fac) * (1024LL * p.ch_fac) + (u % p.ch_fac)));
  i32 *mat = p.matrix + u * p.mat_stride;
  int pos = p.pos[2 * u], f1 = p.pos[2 * u + 1], f2 = f1 + 64;
  const int usb = FRONT ? usb_in : p.usb[u];
  i32 hr_mask = 0;
  {  // ring: HBM -> smem (same layout)
    const i32 *src = reinterpret_cast<const i32 *>(p.states + u * 320);
    i32 *dst = reinterpret_cast<i32 *>(ws.ring);
    for (int i = lane; i < 160; i += 32) dst[i] = __ldg(src + i);
  }
  // lane roles for the grouped modulation
  const int g_slot = lane >> 3;       // slot of the group served in the FFT / post stages
  const int r8 = lane & 7;
  const int fh = r8 >> 2, fi = r8 & 3;  // FFT: half, butterfly position
  __syncwarp();

  // the PCM of the next group of four slots is fetched while this group is processed
  const i32 *w32 = W32 ? p.w32 + u * 1024 : nullptr;
  const int qadj = W32 ? p.qshift_adj[u] : 0;
  // raw loads only (the conversion of a WORD32 sample waits for its load, so it happens where the sample is consumed)
  auto sample = [&](int i) -> i32 { return W32 ? __ldg(w32 + i) : (i32)pcm[(long long)p.ch_fac * i]; };
  i32 nv[4];
#pragma unroll
  for (int s = 0; s < 4; s++) nv[s] = sample(32 * s + lane);
#pragma unroll 1
  for (int g = 0; g < 8; g++) {
    i32 s1v[4], s2v[4];  // fold outputs of the four slots of this group: S1[lane], S2[lane]
    int16_t cv[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
      cv[s] = W32 ? (int16_t)round16(shl32_sat(nv[s], qadj)) : (int16_t)nv[s];
      if (g < 7) nv[s] = sample(32 * (4 * (g + 1) + s) + lane);
    }
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int slot = 4 * g + s;
      // new samples, reversed, into the ring (generic:670-672)
      ws.ring[pos + 31 - lane] = cv[s];
      __syncwarp();
      // 5-tap window (generic:528-588): lane n -> out[n] (fp1, filter_1) and out[32+n] (fp2, filter_2)
      const int16_t *fp1 = ws.ring + ((slot & 1) ? 32 : 0), *fp2 = ws.ring + ((slot & 1) ? 0 : 32);
      i32 a = 0, b = 0;
#pragma unroll
      for (int j = 0; j < 5; j++) {
        a += (i32)fp1[lane + 64 * j] * (i32)sm.qmf_c[f1 + 2 * (lane + 64 * j)];
        b += (i32)fp2[lane + 64 * j] * (i32)sm.qmf_c[f2 + 2 * (lane + 64 * j)];
      }
      __syncwarp();
      pos -= 32;
      if (pos < 0) pos = 288;
      {  // generic:696-718: the two coefficient pointers leap-frog and wrap after 640
        int n1 = f2 + 64, n2 = f1 + 64;
        f1 = n1;
        f2 = n2;
        if (f2 > 640) {
          f1 = 0;
          f2 = 64;
        }
      }
      // generic:480-487: t1 = buf[k] >> 4, t2 = buf[63-k] >> 4
      i32 t1 = a >> 4, t2 = __shfl_sync(full, b, 31 - lane) >> 4;
      s1v[s] = SUB(t1, t2);
      s2v[s] = ADD(t1, t2);
    }
    // ---- pre-twiddle (generic:290-367, M = 16): step n pairs S[n] with S[31-n]; lanes 0-15 serve slots 0/2, 16-31 serve 1/3
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int n = lane & 15;
      // values of slot (2q + (lane>>4)) at positions n and 31-n; odd steps use them swapped (see qmf_synth_kernel.cu)
      const int srcA = (n & 1) ? 31 - n : n, srcB = 31 - srcA;
      i32 a0 = __shfl_sync(full, s1v[2 * q], srcA), a1 = __shfl_sync(full, s1v[2 * q + 1], srcA);
      i32 b0 = __shfl_sync(full, s1v[2 * q], srcB), b1 = __shfl_sync(full, s1v[2 * q + 1], srcB);
      i32 c0 = __shfl_sync(full, s2v[2 * q], srcA), c1 = __shfl_sync(full, s2v[2 * q + 1], srcA);
      i32 d0 = __shfl_sync(full, s2v[2 * q], srcB), d1 = __shfl_sync(full, s2v[2 * q + 1], srcB);
      const bool hi = lane >= 16;
      i32 a = hi ? a1 : a0, b = hi ? b1 : b0, c = hi ? c1 : c0, d = hi ? d1 : d0;
      const int2 tw = sm.pre_tw[n];
      int2 o1, o2;
      o1.x = ADD(__mulhi(a, tw.y), __mulhi(b, tw.x));
      o1.y = SUB(__mulhi(b, tw.y), __mulhi(a, tw.x));
      o2.x = SUB(__mulhi(d, tw.x), __mulhi(c, tw.y));
      o2.y = ADD(__mulhi(c, tw.x), __mulhi(d, tw.y));
      const int e = (n & 1) ? 15 - (n >> 1) : (n >> 1);
      const int sl = 2 * q + (lane >> 4);
      ws.T[sl * kASlot + e] = o1;
      ws.T[sl * kASlot + kAHalf + e] = o2;
    }
    __syncwarp();
    {  // ---- radix-4 stage (generic:1736, index1 = 1, index = 4): legs at fi + 4m ----
      int2 *tb = ws.T + g_slot * kASlot + fh * kAHalf + fi;
      int2 e0 = tb[0], e1 = tb[4], e2 = tb[8], e3 = tb[12];
      const int2 t1 = sm.w[3 * fi], t2 = sm.w[3 * fi + 1], t3 = sm.w[3 * fi + 2];
      i32 xh0 = ADD(e0.x, e2.x), xl0 = SUB(e0.x, e2.x), xh20 = ADD(e1.x, e3.x), xl20 = SUB(e1.x, e3.x);
      i32 xh1 = ADD(e0.y, e2.y), xl1 = SUB(e0.y, e2.y), xh21 = ADD(e1.y, e3.y), xl21 = SUB(e1.y, e3.y);
      i32 xt0 = SUB(xh0, xh20), yt0 = SUB(xh1, xh21);
      i32 xt1 = ADD(xl0, xl21), xt2 = SUB(xl0, xl21);
      i32 yt2 = ADD(xl1, xl20), yt1 = SUB(xl1, xl20);
      tb[0] = make_int2(ADD(xh0, xh20), ADD(xh1, xh21));
      tb[12] = make_int2(lsl(wadd(__mulhi(yt2, t3.x), __mulhi(xt2, t3.y)), 1),
                         lsl(wsub(__mulhi(yt2, t3.y), __mulhi(xt2, t3.x)), 1));
      tb[8] = make_int2(lsl(wadd(__mulhi(yt0, t2.x), __mulhi(xt0, t2.y)), 1),
                        lsl(wsub(__mulhi(yt0, t2.y), __mulhi(xt0, t2.x)), 1));
      tb[4] = make_int2(lsl(wadd(__mulhi(yt1, t1.x), __mulhi(xt1, t1.y)), 1),
                        lsl(wsub(__mulhi(yt1, t1.y), __mulhi(xt1, t1.x)), 1));
    }
    __syncwarp();
    {  // ---- final radix-4 without twiddles + digit reversal (generic:1831-1932): quad fi of half fh ----
      // quad (k, half') = fi: inputs T[4*fi .. 4*fi+3]; outputs F[op], F[op+4], F[op+8], F[op+12], op = digrev[k] + half'
      const int2 *tb = ws.T + g_slot * kASlot + fh * kAHalf + 4 * fi;
      int2 c0 = tb[0], c1 = tb[1], c2 = tb[2], c3 = tb[3];
      i32 xh0 = ADD(c0.x, c2.x), xh1 = ADD(c0.y, c2.y), xl0 = SUB(c0.x, c2.x), xl1 = SUB(c0.y, c2.y);
      i32 zh0 = ADD(c1.x, c3.x), zh1 = ADD(c1.y, c3.y), zl0 = SUB(c1.x, c3.x), zl1 = SUB(c1.y, c3.y);
      int2 *fb = ws.F + g_slot * kASlot + fh * kAHalf + sm.digrev[fi >> 1] + (fi & 1);
      fb[0] = make_int2(ADD(xh0, zh0), ADD(xh1, zh1));
      fb[4] = make_int2(ADD(xl0, zl1), SUB(xl1, zl0));
      fb[8] = make_int2(SUB(xh0, zh0), SUB(xh1, zh1));
      fb[12] = make_int2(SUB(xl0, zl1), ADD(xl1, zl0));
    }
    __syncwarp();
    {  // ---- post-twiddle (generic:388-465, N = 32, H = 8) + t_cos rotation (generic:499-513) + store ----
      const int uu = r8;
      const int2 alt_b = sm.alt_tw[uu], alt_f = sm.alt_tw[uu > 0 ? uu - 1 : 0];
      i32 G1[4], G2[4];  // [0]=G[2u] [1]=G[2u+1] [2]=G[30-2u] [3]=G[31-2u]
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int2 Ff = ws.F[g_slot * kASlot + h * kAHalf + uu], Fb = ws.F[g_slot * kASlot + h * kAHalf + 15 - uu];
        i32 *G = h ? G2 : G1;
        i32 fim = Ff.x, fre = Ff.y;
        i32 t_add = ADD(__mulhi(fre, alt_f.x), __mulhi(fim, alt_f.y));
        i32 t_sub = h ? SUB(__mulhi(fre, alt_f.y), __mulhi(fim, alt_f.x)) : SUB(__mulhi(fim, alt_f.x), __mulhi(fre, alt_f.y));
        if (uu == 0) {
          G[0] = h ? (Ff.y >> 1) : (Ff.x >> 1);
          G[3] = h ? NEG(Ff.x >> 1) : NEG(Ff.y >> 1);
        } else {
          G[0] = h ? t_sub : t_add;
          G[3] = h ? NEG(t_add) : t_sub;
        }
        i32 im = Fb.x, re = Fb.y;
        i32 b_add = ADD(__mulhi(re, alt_b.y), __mulhi(im, alt_b.x));
        i32 b_sub = h ? SUB(__mulhi(re, alt_b.x), __mulhi(im, alt_b.y)) : SUB(__mulhi(im, alt_b.y), __mulhi(re, alt_b.x));
        G[2] = h ? b_sub : b_add;
        G[1] = h ? NEG(b_add) : b_sub;
      }
      const int band[4] = {2 * uu, 2 * uu + 1, 30 - 2 * uu, 31 - 2 * uu};
      i32 ore[4], oim[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        i32 re = G1[j], im = G2[j];
        if (band[j] < usb) {  // always the reference's saturating adds here
          const int2 cs = sm.tcos[band[j]];
          i32 r2 = add_sat(mul32x16_shl(re, cs.x), mul32x16_shl(im, cs.y));
          i32 i2 = sub_sat(mul32x16_shl(im, cs.x), mul32x16_shl(re, cs.y));
          re = r2;
          im = i2;
          if (FRONT) hr_mask |= abs_nrm(re) | abs_nrm(im);
        }
        ore[j] = re;
        oim[j] = im;
      }
      i32 *row = mat + 128 * (4 * g + g_slot);
      *reinterpret_cast<int2 *>(row + 2 * uu) = make_int2(ore[0], ore[1]);
      *reinterpret_cast<int2 *>(row + 30 - 2 * uu) = make_int2(ore[2], ore[3]);
      *reinterpret_cast<int2 *>(row + 64 + 2 * uu) = make_int2(oim[0], oim[1]);
      *reinterpret_cast<int2 *>(row + 64 + 30 - 2 * uu) = make_int2(oim[2], oim[3]);
    }
    __syncwarp();
  }
  {  // ring back to HBM
    i32 *dst = reinterpret_cast<i32 *>(p.states + u * 320);
    const i32 *src = reinterpret_cast<const i32 *>(ws.ring);
    for (int i = lane; i < 160; i += 32) dst[i] = src[i];
    if (lane == 0) {
      p.pos[2 * u] = (int16_t)pos;
      p.pos[2 * u + 1] = (int16_t)f1;
    }
  }
  __syncwarp();
  return hr_mask;
}

template <bool W32>
__global__ void __launch_bounds__(kAnaWarps * 32)
qmf_anal_hq_kernel(QmfAnalArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AnaBlockSmem &sm = *reinterpret_cast<AnaBlockSmem *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const i32 *src = reinterpret_cast<const i32 *>(p.rom);
    i32 *dst = reinterpret_cast<i32 *>(&sm);
    const int nwords = (int)(offsetof(AnaBlockSmem, w_) / 4);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int warps_total = gridDim.x * kAnaWarps;
  for (long long u = (long long)blockIdx.x * kAnaWarps + warp; u < p.n_units; u += warps_total) {
    if (u + warps_total < p.n_units) {  // pull this warp's next unit (2 KB of PCM, 640 B of ring) towards L2
      const long long un = u + warps_total;
      const int16_t *q0 = p.pcm + (p.pcm_unit_stride ? un * p.pcm_unit_stride
                                                       : ((p.ch_fac == 1) ? un * 1024 : (un / p.ch_fac) * (1024LL * p.ch_fac)));
      const int lines = 16 * (p.pcm_unit_stride ? 1 : p.ch_fac);
      if (W32) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(p.w32 + un * 1024) + lane * 128));
      else if (lane < lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(q0) + lane * 128));
      if (lane < 5) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(p.states + un * 320) + lane * 128));
    }
    if (p.exact)
      anal_unit<true, false, W32>(p, sm, sm.w_[warp], u, lane, 0);
    else
      anal_unit<false, false, W32>(p, sm, sm.w_[warp], u, lane, 0);
  }
}

// The front of the fixed-point HQ SBR stage of one unit by one warp (ixheaacd_sbr_dec, decoder/ixheaacd_sbr_dec.c:749-774,
// :790-826, :1050-1127): overlap rows + ixheaacd_rescale_x_overlap, the analysis bank, then the headroom scans / rescale /
// clear of the low band and the HF generator's argument record.  Replaces the sbr_pre_kernel -> qmf_anal_hq_kernel ->
// sbr_scale_kernel sequence: two launches less, the current rows' headroom is collected while they are produced, and
// the rescale pass finds the rows it has just written in L2 (one DRAM write-back per row instead of write + read + write).
template <bool W32>
__global__ void __launch_bounds__(kAnaWarps * 32, 4)
sbr_front_hq_kernel(QmfAnalArgs p, SbrStageArgs g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AnaBlockSmem &sm = *reinterpret_cast<AnaBlockSmem *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const i32 *src = reinterpret_cast<const i32 *>(p.rom);
    i32 *dst = reinterpret_cast<i32 *>(&sm);
    const int nwords = (int)(offsetof(AnaBlockSmem, w_) / 4);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int warps_total = gridDim.x * kAnaWarps;
  for (long long u = (long long)blockIdx.x * kAnaWarps + warp; u < p.n_units; u += warps_total) {
    if (u + warps_total < p.n_units) {  // pull this warp's next unit (PCM, ring, overlap rows, LPC rows) towards L2
      const long long un = u + warps_total;
      const int16_t *q0 = p.pcm + (p.pcm_unit_stride ? un * p.pcm_unit_stride
                                                       : ((p.ch_fac == 1) ? un * 1024 : (un / p.ch_fac) * (1024LL * p.ch_fac)));
      const int lines = 16 * (p.pcm_unit_stride ? 1 : p.ch_fac);
      if (W32) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(p.w32 + un * 1024) + lane * 128));
      else if (lane < lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(q0) + lane * 128));
      if (lane < 5) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(p.states + un * 320) + lane * 128));
      if (lane < 24) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(g.ov + un * 768) + lane * 128));
      if (lane < 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(g.lpc + un * 256) + lane * 128));
    }
    const int usb = sbr_pre_unit(g, u, lane);
    const i32 mask = p.exact ? anal_unit<true, true, W32>(p, sm, sm.w_[warp], u, lane, usb)
                             : anal_unit<false, true, W32>(p, sm, sm.w_[warp], u, lane, usb);
    sbr_scale_unit(g, u, lane, usb, usb <= 32 ? mask : -1);
    __syncwarp();
  }
}

size_t qmf_anal_table_bytes() { return offsetof(AnaBlockSmem, w_); }

// Builds the block-shared table image. Returns 0 if the wrapping path is provably exact for every possible input,
// 1 if the saturating path must be used, -1 on malformed tables.
int qmf_anal_build_tables(const uint8_t *qrom, uint8_t *out) {
  AnaBlockSmem *t = reinterpret_cast<AnaBlockSmem *>(out);
  const int16_t *w16 = reinterpret_cast<const int16_t *>(qrom + kQRomW16);
  const int32_t *dr = reinterpret_cast<const int32_t *>(qrom + kQRomDigRev4_16);
  const int16_t *sc = reinterpret_cast<const int16_t *>(qrom + kQRomSinCosL32);
  const int16_t *al = reinterpret_cast<const int16_t *>(qrom + kQRomAltSinL32);
  const int16_t *tc = reinterpret_cast<const int16_t *>(qrom + kQRomTCosSinL32);
  const int16_t *c = reinterpret_cast<const int16_t *>(qrom + kQRomQmfC);
  auto hi = [](int16_t v) { return (int32_t)((uint32_t)(uint16_t)v << 16); };
  for (int i = 0; i < 1280; i++) t->qmf_c[i] = c[i];
  for (int n = 0; n < 16; n++) t->pre_tw[n] = make_int2(hi(sc[2 * n]), hi(sc[2 * n + 1]));
  for (int n = 0; n < 8; n++) t->alt_tw[n] = make_int2(hi(al[2 * n]), hi(al[2 * n + 1]));
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 3; j++) t->w[3 * i + j] = make_int2(hi(w16[6 * i + 2 * j]), hi(w16[6 * i + 2 * j + 1]));
  for (int k = 0; k < 32; k++) t->tcos[k] = make_int2((int32_t)tc[2 * k], (int32_t)tc[2 * k + 1]);
  for (int k = 0; k < 2; k++) {
    int op = (dr[k] >> 2) >> 1;  // word offset h2 -> complex offset
    if (op < 0 || op + 1 + 12 >= 16) return -1;
    t->digrev[k] = op;
  }
  if (t->digrev[0] == t->digrev[1]) return -1;
  // bounds (see file header)
  long long worst = 0;
  for (int base = 0; base <= 704; base += 64)
    for (int n = 0; n < 32; n++) {
      if (base + 2 * (n + 256) >= 1280) continue;
      long long s = 0;
      for (int j = 0; j < 5; j++) {
        long long v = c[base + 2 * (n + 64 * j)];
        s += v < 0 ? -v : v;
      }
      if (s > worst) worst = s;
    }
  double win = (double)worst * 32768.0;
  if (win >= 2147483647.0) return 1;
  auto smax = [](const int16_t *tab, int pairs) {
    long long m = 0;
    for (int i = 0; i < pairs; i++) {
      long long a = tab[2 * i] < 0 ? -(long long)tab[2 * i] : tab[2 * i];
      long long b = tab[2 * i + 1] < 0 ? -(long long)tab[2 * i + 1] : tab[2 * i + 1];
      if (a + b > m) m = a + b;
    }
    return (double)m;
  };
  double A = 2.0 * (win / 16.0 + 1.0);                       // fold of two >>4 values
  double B = A * smax(sc, 16) / 65536.0 + 2.0;               // pre-twiddle
  double sum4 = 4.0 * B, tw = 2.0 * (sum4 * smax(w16, 12) / 65536.0 + 2.0);
  B = sum4 > tw ? sum4 : tw;                                 // radix-4 stage
  double F = 4.0 * B;                                        // final radix-4
  double G = F * smax(al, 8) / 65536.0 + 2.0;                // post-twiddle
  if (F / 2.0 + 1.0 > G) G = F / 2.0 + 1.0;
  return (G * 1.0001 + 16.0 < 2147483647.0) ? 0 : 1;
}

cudaError_t launch_sbr_front_hq(const QmfAnalArgs &args, const SbrStageArgs &g, int num_sms, cudaStream_t stream) {
  static xb::PerDeviceOnce configured;
  size_t smem = sizeof(AnaBlockSmem);
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(sbr_front_hq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(sbr_front_hq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.done();
  }
  int blocks_per_sm = 4;
  long long need = (args.n_units + kAnaWarps - 1) / kAnaWarps;
  long long grid = (long long)num_sms * blocks_per_sm;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  if (args.w32)
    sbr_front_hq_kernel<true><<<(unsigned)grid, kAnaWarps * 32, smem, stream>>>(args, g);
  else
    sbr_front_hq_kernel<false><<<(unsigned)grid, kAnaWarps * 32, smem, stream>>>(args, g);
  return cudaGetLastError();
}

cudaError_t launch_qmf_anal_hq(const QmfAnalArgs &args, int num_sms, cudaStream_t stream) {
  static xb::PerDeviceOnce configured;
  size_t smem = sizeof(AnaBlockSmem);
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(qmf_anal_hq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(qmf_anal_hq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.done();
  }
  int blocks_per_sm = 4;
  long long need = (args.n_units + kAnaWarps - 1) / kAnaWarps;
  long long grid = (long long)num_sms * blocks_per_sm;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  if (args.w32)
    qmf_anal_hq_kernel<true><<<(unsigned)grid, kAnaWarps * 32, smem, stream>>>(args);
  else
    qmf_anal_hq_kernel<false><<<(unsigned)grid, kAnaWarps * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace xb
