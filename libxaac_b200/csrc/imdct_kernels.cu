// imdct_kernels.cu — AAC 1024/128 IMDCT + window/overlap-add for sm_100a (B200).
//
// One warp owns one unit (frame x channel) end to end; a persistent grid walks the batch.
// Replaces, bit-exactly, the reference stage ixheaacd_imdct_process (decoder/ixheaacd_lpfuncs.c:347-802)
// and the leaves it dispatches to through ixheaacd_function_selector.h:
//   ixheaacd_calc_max_spectral_line_dec   decoder/ixheaacd_aac_tns.c:422
//   ixheaacd_pretwiddle_compute_dec       decoder/ixheaacd_aac_imdct.c:165
//   ixheaacd_imdct_using_fft_dec          decoder/ixheaacd_aac_imdct.c:834   (radix-8 DIT, wrapping)
//   ixheaacd_post_twiddle_dec             decoder/ixheaacd_aac_imdct.c:331
//   ixheaacd_post_twid_overlap_add_dec    decoder/ixheaacd_aac_imdct.c:506   (long->long fused)
//   ixheaacd_process_win_seq / long_short_win_seq / over_lap_add1/2 / spec_to_overlapbuf ...
//                                         decoder/ixheaacd_lpfuncs.c:94-345, decoder/ixheaacd_block.c:1193-1240
//
// Data flow per unit (long block):
//   HBM spec[1024] --LDG.64, coalesced--> registers (16 int2 / lane)
//     -> warp OR-reduce (block exponent) -> fold + pre-twiddle (partner sample via one SHFL, lane^31)
//     -> smem X[512 complex] -> radix-8 stage 1 -> smem Y (XOR-swizzled, conflict-free for all 3 stages)
//     -> stage 2, stage 3 in place -> post-twiddle fused with sine/KBD window + overlap-add
//   HBM overlap[512] read once (LDG.64) and written once (STG.64); WORD32 out[1024] written once.
// The butterfly graph, operation order and truncation points are those of the reference;
// shuffles/shared memory only move data.
#include <cstdint>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"

namespace xb {

constexpr int kWarpsPerBlock = 8;
constexpr int kSeqOnlyLong = 0, kSeqLongStart = 1, kSeqEightShort = 2, kSeqLongStop = 3;

struct WarpSmem {
  int2 X[512];  // pre-twiddle output / later the post-twiddled real block T[1024]
  int2 Y[512];  // FFT workspace (swizzled for the 512-point transform) / later overlap staging
};

struct BlockSmem {
  i32 cs[257];   // cosine_array_2048_256 as packed pairs: lo16 = A_p, hi16 = B_p
  i32 tw[448];   // fft_twiddle
  WarpSmem w[kWarpsPerBlock];
};

// physical slot of complex element i of the 512-point workspace. i = 64k + 8q + m (octal digits):
// stage 1 writes vary (k, q&1), stage 2 varies (m|k pairs), stage 3 and the post-twiddle vary (m, q&1);
// this XOR keeps every half-warp 64-bit access on 16 distinct bank pairs.
XB_DEV int swz(int i) {
  int k = i >> 6;
  return (i & ~15) | ((i ^ k) & 7) | ((((i >> 3) ^ k) & 1) << 3);
}

struct C8 {
  i32 r[8], i[8];
};

// Radix-8 butterfly core, aac_imdct.c:876-999 (k1 = 1) and :1213-1375 (k1 = 2: legs 1,2,4,6 arrive
// doubled after the twiddle multiply, legs 3,5,7 do not). Results are left in storage order.
// k1 is a multiplier rather than a shift count so that a run-time value costs one IMAD (fma pipe), not SHF+IADD.
XB_DEV void bfly8(C8 &x, const i32 k1) {
  i32 t;
  const i32 k2 = k1 * 2;
#define SH(v, s) lsl((v), (s))
#define M1(v) ((i32)((u32)(v) * (u32)k1))
#define M2(v) ((i32)((u32)(v) * (u32)k2))
  x.r[0] = wadd(x.r[0], x.r[4]); x.i[0] = wadd(x.i[0], x.i[4]);
  x.r[4] = wsub(x.r[0], SH(x.r[4], 1)); x.i[4] = wsub(x.i[0], SH(x.i[4], 1));
  x.r[2] = wadd(x.r[2], x.r[6]); x.i[2] = wadd(x.i[2], x.i[6]);
  x.r[6] = wsub(x.r[2], SH(x.r[6], 1)); x.i[6] = wsub(x.i[2], SH(x.i[6], 1));
  x.r[0] = wadd(x.r[0], x.r[2]); x.i[0] = wadd(x.i[0], x.i[2]);
  x.r[2] = wsub(x.r[0], SH(x.r[2], 1)); x.i[2] = wsub(x.i[0], SH(x.i[2], 1));
  x.r[4] = wadd(x.r[4], x.i[6]); x.i[4] = wsub(x.i[4], x.r[6]);
  t = x.r[6];
  x.r[6] = wsub(x.r[4], SH(x.i[6], 1)); x.i[6] = wadd(x.i[4], SH(t, 1));

  x.r[1] = wadd(x.r[1], M1(x.r[5])); x.i[1] = wadd(x.i[1], M1(x.i[5]));
  x.r[5] = wsub(x.r[1], M2(x.r[5])); x.i[5] = wsub(x.i[1], M2(x.i[5]));
  x.r[3] = wadd(x.r[3], x.r[7]); x.i[3] = wadd(x.i[3], x.i[7]);
  x.r[7] = wsub(x.r[3], SH(x.r[7], 1)); x.i[7] = wsub(x.i[3], SH(x.i[7], 1));
  x.r[1] = wadd(x.r[1], M1(x.r[3])); x.i[1] = wadd(x.i[1], M1(x.i[3]));
  x.r[3] = wsub(x.r[1], M2(x.r[3])); x.i[3] = wsub(x.i[1], M2(x.i[3]));
  x.r[5] = wadd(x.r[5], x.i[5]); x.i[5] = wsub(x.r[5], SH(x.i[5], 1));
  x.r[7] = wadd(x.r[7], x.i[7]); x.i[7] = wsub(x.r[7], SH(x.i[7], 1));
  x.i[7] = wsub(x.r[5], M1(x.i[7])); x.r[5] = wsub(x.i[7], SH(x.r[5], 1));
  x.i[5] = wsub(M1(x.r[7]), x.i[5]); x.r[7] = wsub(x.i[5], M2(x.r[7]));
  x.i[7] = SH(x.i[7], 1); x.r[5] = SH(x.r[5], 1); x.i[5] = SH(x.i[5], 1); x.r[7] = SH(x.r[7], 1);

  x.r[0] = wadd(x.r[0], x.r[1]); x.i[0] = wadd(x.i[0], x.i[1]);
  x.r[1] = wsub(x.r[0], SH(x.r[1], 1)); x.i[1] = wsub(x.i[0], SH(x.i[1], 1));
  x.r[2] = wadd(x.r[2], x.i[3]);
  t = wsub(x.r[2], SH(x.i[3], 1));
  x.i[2] = wsub(x.i[2], x.r[3]);
  x.i[3] = wadd(x.i[2], SH(x.r[3], 1));
  const i32 k = 0x5A82 << 16;
  i32 p7i = wadd(x.r[4], __mulhi(x.i[7], k)); i32 n4r = wsub(p7i, SH(x.r[4], 1));
  i32 p7r = wadd(x.i[4], __mulhi(x.r[7], k)); i32 n4i = wsub(p7r, SH(x.i[4], 1));
  i32 p5i = wadd(x.r[6], __mulhi(x.i[5], k)); i32 n6r = wsub(p5i, SH(x.r[6], 1));
  i32 p5r = wadd(x.i[6], __mulhi(x.r[5], k)); i32 n6i = wsub(p5r, SH(x.i[6], 1));
  // storage order: 0:x0 1:(x7i,x7r) 2:x2 3:(x5i,x5r) 4:x1 5:-x4 6:(t,x3i) 7:-x6
  i32 r1 = x.r[1], i1 = x.i[1], i3 = x.i[3];
  x.r[1] = p7i; x.i[1] = p7r;
  x.r[3] = p5i; x.i[3] = p5r;
  x.r[4] = r1;  x.i[4] = i1;
  x.r[5] = wneg(n4r); x.i[5] = wneg(n4i);
  x.r[6] = t;   x.i[6] = i3;
  x.r[7] = wneg(n6r); x.i[7] = wneg(n6i);
#undef SH
#undef M1
#undef M2
}

// aac_imdct.c:1179-1185 (doubled) / :1256-1260 (plain)
XB_DEV void tw_mul(i32 &re, i32 &im, i32 w, int dbl) {
  i32 a = wsub(mul32x16l(re, w), mul32x16h(im, w));
  i32 b = wadd(mul32x16h(re, w), mul32x16l(im, w));
  re = lsl(a, dbl);
  im = lsl(b, dbl);
}

XB_DEV void tw_all(C8 &x, const i32 *tw, int step) {
#pragma unroll
  for (int q = 1; q < 8; q++) tw_mul(x.r[q], x.i[q], tw[q * step], (q == 1 || !(q & 1)) ? 1 : 0);
}

template <bool SWZ>
XB_DEV int2 ldY(const int2 *Y, int i) { return Y[SWZ ? swz(i) : i]; }
template <bool SWZ>
XB_DEV void stY(int2 *Y, int i, i32 r, i32 im) { Y[SWZ ? swz(i) : i] = make_int2(r, im); }

// (C,S) selection shared by pre- and post-twiddle (see oracle/src/imdct.c cs_pair for the derivation
// from aac_imdct.c:176-239 and :346-404). c = complex bin, half = n/4, words-per-pair stride wst.
XB_DEV void cs_pair(const i32 *cs, int c, int quarter, int wst, i32 &C, i32 &S) {
  if (c <= quarter) {
    i32 w = cs[wst * c];
    C = (i32)((u32)w << 16);
    S = (i32)((u32)w & 0xffff0000u);
  } else {
    i32 w = cs[wst * (2 * quarter - c)];
    S = (i32)((u32)w << 16);
    C = (i32)((u32)w & 0xffff0000u);
  }
}

// post-twiddle of one bin: aac_imdct.c:351-362 (adjust = +-50 long, +-402 short)
XB_DEV void post_bin(int2 y, i32 C, i32 S, i32 adj_hi, i32 &outr, i32 &outi) {
  i32 orr = wadd(__mulhi(y.x, C), __mulhi(y.y, S));
  i32 oi = wsub(__mulhi(y.x, S), __mulhi(y.y, C));
  outr = wadd(orr, __mulhi(oi, wneg(adj_hi)));
  outi = wadd(oi, __mulhi(orr, adj_hi));
}

// ---- the rare-path helpers below work on T (post-twiddled block, smem) and P (overlap copy, smem) ----

// block.c:1193-1218
__device__ __noinline__ void ola1(const i32 *coef, const i32 *prev, i32 *out, const i16 *w, int q_shift, int size, int ch_fac,
                 int lane) {
  for (int i = lane; i < size; i += 32) {
    i32 w1 = w[2 * size - 2 * i - 1], w2 = w[2 * size - 2 * i - 2];
    i32 c = coef[2 * size - 1 - i];
    out[ch_fac * (size - 1 - i)] =
        sub_sat(shl32_dir_sat_limit(mul32x16(c, w2), q_shift), mul32x16_fullsat(prev[i], w1));
    out[ch_fac * (size + i)] =
        sub_sat(shl32_dir_sat_limit(mul32x16(neg_sat(c), w1), q_shift), mul32x16_fullsat(prev[i], w2));
  }
}

// block.c:1220-1240 (ch_fac == 1 at every call site). prev/out may alias different parts of P.
__device__ __noinline__ void ola2(const i32 *coef, const i32 *prev, i32 *out, const i16 *w, int q_shift, int size, int lane) {
  for (int i = lane; i < size; i += 32) {
    i32 a = sub_sat(mul32x16(coef[size + i], w[2 * i]), mul32x16(prev[size - 1 - i], w[2 * i + 1]));
    i32 b = sub_sat(mul32x16(neg_sat(coef[2 * size - 1 - i]), w[2 * size - 2 * i - 1]),
                    mul32x16(prev[i], w[2 * size - 2 * i - 2]));
    out[i] = shr32_sat(a, 16 - (q_shift + 1));
    out[i + size] = shr32_sat(b, 16 - (q_shift + 1));
  }
}

// lpfuncs.c:94-178
__device__ __noinline__ void process_win_seq(const i32 *coef, const i32 *prev, i32 *out, const i16 *wl, const i16 *ws, int q_shift,
                            int ch_fac, int flag, int lane) {
  const int s1 = 64, s7 = 448, s8 = 512, s9 = 576, s14 = 896, s15 = 960;
  const i16 *w_sh, *w_lg;
  const i32 *pv;
  if (flag) {
    for (int i = lane; i < s7; i += 32) {
      i32 t = shl32_dir_sat_limit(mul32x16(coef[s8 + i], wl[2 * i]), q_shift + 1);
      out[ch_fac * i] = add_sat(t, lsl(prev[i], 16));
      i32 a = shl32_dir_sat_limit(mul32x16(wneg(coef[s15 - 1 - i]), wl[2 * (s7 - i) - 1]), q_shift);
      out[ch_fac * (i + s9)] = lsl(a, 1);
    }
    w_sh = ws;
    w_lg = wl + s14;
    pv = prev + s8 - 1;
  } else {
    for (int i = lane; i < s7; i += 32) {
      out[ch_fac * i] = mul32x16_fullsat(prev[s8 - 1 - i], neg16(wl[2 * i + 1]));
      out[ch_fac * (s9 + i)] = sub_sat(shl32_dir_sat_limit(wneg(coef[s15 - 1 - i]), q_shift - 1),
                                       mul32x16_fullsat(prev[i + s1], wl[2 * s7 - 2 - 2 * i]));
    }
    w_sh = wl + s14;
    w_lg = ws;
    pv = prev + s1 - 1;
  }
  for (int k = lane; k < s1; k += 32) {
    i32 c = coef[s15 + k];
    i32 win1 = w_lg[2 * k], win2 = w_lg[2 * k + 1];
    i32 win4 = w_sh[2 * k], win3 = w_sh[2 * k + 1];
    i32 p = pv[-k];
    i32 a = sub_sat(shl32_dir_sat_limit(mul32x16(c, win1), q_shift), mul32x16_fullsat(p, win3));
    out[ch_fac * (s7 + k)] = lsl(a, flag);
    a = sub_sat(shl32_dir_sat_limit(mul32x16(neg_sat(c), win2), q_shift), mul32x16_fullsat(p, win4));
    out[ch_fac * (s9 - 1 - k)] = lsl(a, flag);
  }
}

// lpfuncs.c:218-284 incl. the four long_short_win_process calls (:180-216)
__device__ __noinline__ void long_short_win_seq(const i32 *cur, i32 *prev, i32 *out, const i16 *sw, const i16 *swp, const i16 *lwp,
                               int q_shift, int ch_fac, int lane) {
  const int s1 = 64, s2 = 128, s3 = 192, s6 = 384, s7 = 448, s8 = 512, s9 = 576, s10 = 640, s16 = 1024;
  for (int i = lane; i < s7; i += 32) out[ch_fac * i] = mul32x16_fullsat(prev[s8 - 1 - i], neg16(lwp[2 * i + 1]));
  for (int i = lane; i < s1; i += 32) {
    out[ch_fac * (s7 + i)] = sub_sat(shl32_dir_sat_limit(mul32x16(cur[s1 + i], swp[2 * i]), q_shift),
                                     mul32x16_fullsat(prev[s1 - 1 - i], lwp[2 * s7 + 1 + 2 * i]));
    out[ch_fac * (s8 + i)] =
        sub_sat(shl32_dir_sat_limit(mul32x16(neg_sat(cur[s2 - 1 - i]), swp[s2 - 2 * i - 1]), q_shift),
                mul32x16_fullsat(prev[i], lwp[s16 - 2 - 2 * i]));
  }
  for (int b = 0; b < 4; b++) {
    int inc = b * s2;
    const i32 *c0 = cur + s1 + inc;
    const i32 *p0 = prev + s1 + inc;
    i32 *o0 = out + ch_fac * (s9 + inc);
    const i16 *lw = lwp + 2 * (s7 - inc);
    for (int i = lane; i < s1; i += 32) {
      int j = s1 - 1 - i;
      i32 c1 = c0[s3 - 1 - j], c2 = c0[-s1 + j];
      i32 sh1 = sw[s2 - 1 - 2 * j], sh2 = sw[s2 - 2 - 2 * j];
      o0[ch_fac * i] = sub_sat(shl32_dir_sat_limit(wsub(mul32x16(c1, sh2), mul32x16(c2, sh1)), q_shift),
                               mul32x16_fullsat(p0[i], lw[-2 - 2 * i]));
      if (b != 3)
        o0[ch_fac * (s2 - 1 - i)] =
            sub_sat(shl32_dir_sat_limit(wsub(mul32x16(neg_sat(c1), sh1), mul32x16(c2, sh2)), q_shift),
                    mul32x16_fullsat(p0[s2 - 1 - i], lw[-2 * s2 + 2 * i]));
    }
  }
  __syncwarp();
  for (int i = lane; i < s1; i += 32) {
    i32 a = wsub(mul32x16(wneg(cur[s10 - 1 - i]), sw[s2 - 2 * i - 1]), mul32x16(cur[s6 + i], sw[s2 - 2 * i - 2]));
    prev[i] = round16(shl32_dir_sat_limit(a, q_shift + 1));
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32, 3)
imdct_ola_kernel(ImdctArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BlockSmem &sm = *reinterpret_cast<BlockSmem *>(smem_raw);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;

  {  // block-shared ROM: cos/sin pairs and FFT twiddles
    const i32 *cs_g = reinterpret_cast<const i32 *>(p.rom + kDevCos);
    const i32 *tw_g = reinterpret_cast<const i32 *>(p.rom + kDevFftTw);
    for (int i = threadIdx.x; i < 257; i += blockDim.x) sm.cs[i] = cs_g[i];
    for (int i = threadIdx.x; i < 448; i += blockDim.x) sm.tw[i] = tw_g[i];
  }
  __syncthreads();
  // window tables stay in global memory (L1/L2 resident, read-only path); shape 0 = sine, 1 = KBD
  auto win_long = [&](int shape) {
    return reinterpret_cast<const i16 *>(p.rom + kDevWinLongSine + shape * (kDevWinLongKbd - kDevWinLongSine));
  };
  auto win_short = [&](int shape) {
    return reinterpret_cast<const i16 *>(p.rom + kDevWinShortSine + shape * (kDevWinShortKbd - kDevWinShortSine));
  };

  int2 *X = sm.w[warp].X;
  int2 *Y = sm.w[warp].Y;
  const int warps_total = gridDim.x * kWarpsPerBlock;

  for (long long u = (long long)blockIdx.x * kWarpsPerBlock + warp; u < p.n_units; u += warps_total) {
    const int2 *spec2 = reinterpret_cast<const int2 *>(p.spec + u * 1024);
    i32 *ovl_g = p.overlap + u * 512;
    const int ch_fac = p.ch_fac;
    // ch_fac == 1: planar, unit-major. ch_fac > 1: consecutive units are the channels of one frame and are
    // interleaved sample-wise exactly like the reference's time buffer (out_samples + ch, stride ch_fac).
    i32 *out_g = (ch_fac == 1) ? p.out + u * 1024 : p.out + (u / ch_fac) * (1024LL * ch_fac) + (u % ch_fac);
    const int win_seq = p.ics[2 * u], win_shape = p.ics[2 * u + 1];
    const int prev_shape = p.wstate[2 * u], prev_seq = p.wstate[2 * u + 1];
    const bool prev_longish = (prev_seq == kSeqOnlyLong) || (prev_seq == kSeqLongStop);

    // ---- load + block exponent (aac_tns.c:422) ----
    int2 v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = __ldg(spec2 + lane + 32 * k);
    i32 acc = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) acc |= abs_nrm(v[k].x) | abs_nrm(v[k].y);
    acc = __reduce_or_sync(full, acc);
    const int headroom = norm32(acc);
    int q_shift, adj;

    if (win_seq != kSeqEightShort) {
      const int expo = 8 - (headroom - 1);  // lpfuncs.c:414-416
      // ---- fold + pre-twiddle (aac_imdct.c:165-329) ----
#pragma unroll
      for (int k = 0; k < 16; k++) {
        int c = lane + 32 * k;
        i32 xr = v[k].x;
        i32 xi = __shfl_sync(full, v[15 - k].y, lane ^ 31);
        i32 C, S;
        cs_pair(sm.cs, c, 256, 1, C, S);
        i32 re = wadd(__mulhi(xr, C), __mulhi(xi, S));
        i32 im = wsub(__mulhi(xi, C), __mulhi(xr, S));
        if (expo < 0) {
          re = shl32(re, -expo);
          im = shl32(im, -expo);
        } else {
          re = shr32(re, expo);
          im = shr32(im, expo);
        }
        X[c] = make_int2(re, im);
      }
      __syncwarp();
      // ---- radix-8 stage 1 (aac_imdct.c:856-1000): digit-reversed gather == stride-64 read ----
      // The two butterflies of a lane are a rolled loop on purpose: the unrolled body overflowed the 32 KB
      // instruction cache (ncu: stall_no_instruction dominant, profiles/r1_imdct_a.md).
      // Output slot of leg p: swz(64*(b&7) + 8*(b>>3) + p) == s1base ^ p  (p < 8 only touches the low 3 bits).
#pragma unroll 1
      for (int t = 0; t < 2; t++) {
        const int b = lane + 32 * t;
        C8 x;
#pragma unroll
        for (int q = 0; q < 8; q++) {
          int2 e = X[b + 64 * q];
          x.r[q] = e.x;
          x.i[q] = e.y;
        }
        bfly8(x, 1);
        const int kk = b & 7;
        const int s1base = (kk << 6) | ((b >> 4) << 4) | ((((b >> 3) ^ kk) & 1) << 3) | kk;
#pragma unroll
        for (int q = 0; q < 8; q++) Y[s1base ^ q] = make_int2(x.r[q], x.i[q]);
      }
      __syncwarp();
      // ---- stage 2 (del = 8): 56 twiddled columns + 8 plain ones (aac_imdct.c:1007-1384) ----
      // column m, block k: element q sits at swz(m + 64k + 8q) == (s2base ^ ((q&1)<<3)) + 16*(q>>1)
#pragma unroll 1
      for (int t = 0; t < 2; t++) {
        const int k = lane & 7;
        const int m = (t == 0) ? 1 + (lane >> 3) : (lane < 24 ? 5 + (lane >> 3) : 0);
        const int s2base = (k << 6) | ((k & 1) << 3) | ((m ^ k) & 7);
        int2 *Ye = Y + s2base;
        int2 *Yo = Y + (s2base ^ 8);
        C8 x;
#pragma unroll
        for (int q = 0; q < 8; q++) {
          int2 e = ((q & 1) ? Yo : Ye)[16 * (q >> 1)];
          x.r[q] = e.x;
          x.i[q] = e.y;
        }
        if (m != 0) tw_all(x, sm.tw, 8 * m);
        bfly8(x, m != 0 ? 2 : 1);
#pragma unroll
        for (int q = 0; q < 8; q++) ((q & 1) ? Yo : Ye)[16 * (q >> 1)] = make_int2(x.r[q], x.i[q]);
      }
      __syncwarp();
      // ---- stage 3 (del = 64), all columns twiddled incl. column 0 (aac_imdct.c:1386-1621) ----
      // element q of column m sits at swz(m + 64q) == 64q + (m ^ c_q), c_q = (q&7) | ((q&1)<<3)
#pragma unroll 1
      for (int t = 0; t < 2; t++) {
        const int m = lane + 32 * t;
        C8 x;
#pragma unroll
        for (int q = 0; q < 8; q++) {
          int2 e = Y[64 * q + (m ^ ((q & 7) | ((q & 1) << 3)))];
          x.r[q] = e.x;
          x.i[q] = e.y;
        }
        tw_all(x, sm.tw, m);
        bfly8(x, 2);
#pragma unroll
        for (int q = 0; q < 8; q++) Y[64 * q + (m ^ ((q & 7) | ((q & 1) << 3)))] = make_int2(x.r[q], x.i[q]);
      }
      __syncwarp();
      q_shift = (31 + expo + 2) - 26;  // lpfuncs.c:423 with imdct_scale = expo + 2

      const i32 adj_hi = 50 << 16;
      if (win_seq == kSeqOnlyLong && prev_longish) {
        // ---- fused post-twiddle + window + OLA (aac_imdct.c:506-832) ----
        const i16 *win = win_long(prev_shape);
        const int2 *win4 = reinterpret_cast<const int2 *>(win);
        int2 *ovl2 = reinterpret_cast<int2 *>(ovl_g);
#pragma unroll 2
        for (int j = 0; j < 8; j++) {
          int c = lane + 32 * j, c2 = 511 - c;
          i32 C, S, r1, i1, r2, i2;
          cs_pair(sm.cs, c, 256, 1, C, S);
          post_bin(ldY<true>(Y, c), C, S, adj_hi, r1, i1);
          cs_pair(sm.cs, c2, 256, 1, C, S);
          post_bin(ldY<true>(Y, c2), C, S, adj_hi, r2, i2);
          int2 pv = ovl2[c];
          int2 wq = __ldg(win4 + (255 - c));  // win[1020-4c .. 1023-4c]
          i32 wlo_b = sext16(wq.x), whi_b = wq.x >> 16;  // taps for m' = 510-2c
          i32 wlo_a = sext16(wq.y), whi_a = wq.y >> 16;  // taps for m  = 511-2c
          i32 xa = i1, xb = r2, pa = pv.x, pb = pv.y;
          i32 a0, a1, b0, b1;
          if (q_shift > 0) {
            a0 = shl32_sat(mul32x16(xa, wlo_a), q_shift);
            a1 = shl32_sat(mul32x16(neg_sat(xa), whi_a), q_shift);
            b0 = shl32_sat(mul32x16(xb, wlo_b), q_shift);
            b1 = shl32_sat(mul32x16(neg_sat(xb), whi_b), q_shift);
          } else {
            pa = sext16(pa);  // aac_imdct.c:679: overlap read through a WORD16
            pb = sext16(pb);
            a0 = shr32(mul32x16(xa, wlo_a), -q_shift);
            a1 = shr32(mul32x16(neg_sat(xa), whi_a), -q_shift);
            b0 = shr32(mul32x16(xb, wlo_b), -q_shift);
            b1 = shr32(mul32x16(neg_sat(xb), whi_b), -q_shift);
          }
          i32 o_m = sub_sat(a0, mul32x16_fullsat(pa, whi_a));    // out[511-2c]
          i32 o_M = sub_sat(a1, mul32x16_fullsat(pa, wlo_a));    // out[512+2c]
          i32 o_m2 = sub_sat(b0, mul32x16_fullsat(pb, whi_b));   // out[510-2c]
          i32 o_M2 = sub_sat(b1, mul32x16_fullsat(pb, wlo_b));   // out[513+2c]
          ovl2[c] = make_int2(shr32_sat(r1, 16 - q_shift), shr32_sat(i2, 16 - q_shift));
          if (ch_fac == 1) {
            *reinterpret_cast<int2 *>(out_g + 510 - 2 * c) = make_int2(o_m2, o_m);
            *reinterpret_cast<int2 *>(out_g + 512 + 2 * c) = make_int2(o_M, o_M2);
          } else {
            out_g[ch_fac * (510 - 2 * c)] = o_m2;
            out_g[ch_fac * (511 - 2 * c)] = o_m;
            out_g[ch_fac * (512 + 2 * c)] = o_M;
            out_g[ch_fac * (513 + 2 * c)] = o_M2;
          }
        }
        adj = 2;
      } else {
        // ---- un-fused: post-twiddle to T (aac_imdct.c:331-421), then the window-sequence variants ----
        i32 *T = reinterpret_cast<i32 *>(X);
        i32 *P = reinterpret_cast<i32 *>(Y);  // overlap staging (after Y has been consumed)
#pragma unroll 4
        for (int j = 0; j < 16; j++) {
          int c = lane + 32 * j;
          i32 C, S, r1, i1;
          cs_pair(sm.cs, c, 256, 1, C, S);
          post_bin(ldY<true>(Y, c), C, S, adj_hi, r1, i1);
          T[2 * c] = r1;
          T[1023 - 2 * c] = i1;
        }
        __syncwarp();
        for (int i = lane; i < 512; i += 32) P[i] = ovl_g[i];
        __syncwarp();
        const i16 *wl = win_long(prev_shape);
        const i16 *wsp = win_short(prev_shape);
        const int s1 = 64, s7 = 448, s8 = 512, s9 = 576, s14 = 896;
        if (win_seq == kSeqOnlyLong) {  // previous was start/short (lpfuncs.c:489-521)
          process_win_seq(T, P, out_g, wl, wsp, q_shift, ch_fac, 1, lane);
          __syncwarp();
          for (int i = lane; i < s8; i += 32) P[i] = shr32_sat(T[i], 16 - q_shift);
          adj = 1;
        } else if (win_seq == kSeqLongStart) {  // lpfuncs.c:526-581
          if (prev_longish) {
            ola1(T, P, out_g, wl, q_shift, s8, ch_fac, lane);
            adj = 2;
          } else {
            process_win_seq(T, P, out_g, wl, wsp, q_shift, ch_fac, 1, lane);
            adj = 1;
          }
          __syncwarp();
          for (int i = lane; i < s7; i += 32) P[i] = shr32_sat(neg_sat(T[s1 + s7 - 1 - i]), 16 - q_shift);
          for (int i = lane; i < s1; i += 32) P[s7 + i] = shr32_sat(T[i], 16 - q_shift);
        } else {  // LONG_STOP (lpfuncs.c:583-654)
          if (!prev_longish) {
            for (int i = lane; i < s7; i += 32) {
              out_g[ch_fac * i] = shl32_sat(sext16(P[i]), 15);
              out_g[ch_fac * (s9 + i)] = shl32_dir_sat_limit(neg_sat(T[s8 + s7 - 1 - i]), q_shift - 1);
            }
            ola1(T + s14, P + s7, out_g + ch_fac * s7, wsp, q_shift, s1, ch_fac, lane);
          } else {
            process_win_seq(T, P, out_g, wl, wsp, q_shift, ch_fac, 0, lane);
          }
          __syncwarp();
          for (int i = lane; i < s8; i += 32) P[i] = shr32_sat(T[i], 16 - q_shift);
          adj = 2;
        }
        __syncwarp();
        for (int i = lane; i < 512; i += 32) ovl_g[i] = P[i];
      }
    } else {
      // ================= EIGHT_SHORT: 8 x (128-point IMDCT) (lpfuncs.c:657-798) =================
      const int expo = 5 - (headroom - 1);
#pragma unroll
      for (int k = 0; k < 16; k++) {
        int c = lane + 32 * k;       // pair index; window w = c>>6, bin cw = c&63
        int cw = c & 63;
        i32 xr = v[k].x;
        i32 xi = __shfl_sync(full, v[k ^ 1].y, lane ^ 31);
        i32 C, S;
        cs_pair(sm.cs, cw, 32, 8, C, S);
        i32 re = wadd(__mulhi(xr, C), __mulhi(xi, S));
        i32 im = wsub(__mulhi(xi, C), __mulhi(xr, S));
        if (expo < 0) {
          re = shl32(re, -expo);
          im = shl32(im, -expo);
        } else {
          re = shr32(re, expo);
          im = shr32(im, expo);
        }
        X[c] = make_int2(re, im);
      }
      __syncwarp();
      // 64-point FFT per window: stage 1 (identity digit reversal) then the final twiddled stage
#pragma unroll 1
      for (int t = 0; t < 2; t++) {
        int j = lane + 32 * t;
        int wofs = (j >> 3) << 6, g = j & 7;
        C8 x;
#pragma unroll
        for (int q = 0; q < 8; q++) {
          int2 e = X[wofs + g + 8 * q];
          x.r[q] = e.x;
          x.i[q] = e.y;
        }
        bfly8(x, 1);
#pragma unroll
        for (int q = 0; q < 8; q++) stY<false>(Y, wofs + 8 * g + q, x.r[q], x.i[q]);
      }
      __syncwarp();
#pragma unroll 1
      for (int t = 0; t < 2; t++) {
        int j = lane + 32 * t;
        int wofs = (j >> 3) << 6, m = j & 7;
        C8 x;
#pragma unroll
        for (int q = 0; q < 8; q++) {
          int2 e = ldY<false>(Y, wofs + m + 8 * q);
          x.r[q] = e.x;
          x.i[q] = e.y;
        }
        tw_all(x, sm.tw, 8 * m);
        bfly8(x, 2);
#pragma unroll
        for (int q = 0; q < 8; q++) stY<false>(Y, wofs + m + 8 * q, x.r[q], x.i[q]);
      }
      __syncwarp();
      q_shift = (31 + expo + 2) - 23;  // lpfuncs.c:682-683
      i32 *T = reinterpret_cast<i32 *>(X);
      i32 *P = reinterpret_cast<i32 *>(Y);
      const i32 adj_hi = 402 << 16;
#pragma unroll 4
      for (int k = 0; k < 16; k++) {
        int c = lane + 32 * k;
        int cw = c & 63, wbase = (c >> 6) << 7;
        i32 C, S, r1, i1;
        cs_pair(sm.cs, cw, 32, 8, C, S);
        post_bin(ldY<false>(Y, c), C, S, adj_hi, r1, i1);
        T[wbase + 2 * cw] = r1;
        T[wbase + 127 - 2 * cw] = i1;
      }
      __syncwarp();
      for (int i = lane; i < 512; i += 32) P[i] = ovl_g[i];
      __syncwarp();
      const i16 *sw = win_short(win_shape);
      const i16 *wsp = win_short(prev_shape);
      const i16 *wl = win_long(prev_shape);
      const int s1 = 64, s2 = 128, s6 = 384, s7 = 448, s8 = 512, s9 = 576, s10 = 640, s14 = 896, s15 = 960;
      if (!prev_longish) {
        for (int i = lane; i < s7; i += 32) out_g[ch_fac * i] = shl32_sat(sext16(P[i]), 15);
        ola1(T, P + s7, out_g + ch_fac * s7, wsp, q_shift, s1, ch_fac, lane);
        for (int b = 0; b < 3; b++) {
          int inc = b * s2;
          // spec_to_overlapbuf into a local buffer then over_lap_add1 (lpfuncs.c:719-737): fused per element
          for (int i = lane; i < s1; i += 32) {
            i32 pvl = shr32_sat(T[inc + i], 16 - q_shift);
            i32 w1 = sw[2 * s1 - 2 * i - 1], w2 = sw[2 * s1 - 2 * i - 2];
            i32 c = T[s2 + inc + 2 * s1 - 1 - i];
            i32 *o = out_g + ch_fac * (s9 + inc);
            o[ch_fac * (s1 - 1 - i)] =
                sub_sat(shl32_dir_sat_limit(mul32x16(c, w2), q_shift), mul32x16_fullsat(pvl, w1));
            o[ch_fac * (s1 + i)] =
                sub_sat(shl32_dir_sat_limit(mul32x16(neg_sat(c), w1), q_shift), mul32x16_fullsat(pvl, w2));
          }
        }
        __syncwarp();
        ola2(T + s8, T + s6, P, sw, q_shift, s1, lane);
        __syncwarp();
        for (int i = lane; i < s1; i += 32) {  // lpfuncs.c:335-345
          out_g[ch_fac * (s15 + i)] = shl32_sat(sext16(P[i]), 15);
          P[i] = P[s1 + i];
        }
      } else {
        long_short_win_seq(T, P, out_g, sw, wsp, wl, q_shift, ch_fac, lane);
      }
      __syncwarp();
      for (int b = 0; b < 3; b++) {
        int inc = b * s2;
        ola2(T + s10 + inc, T + s8 + inc, P + s1 + inc, sw, q_shift, s1, lane);
      }
      for (int i = lane; i < s1; i += 32) P[s7 + i] = shr32_sat(T[s14 + i], 16 - q_shift);
      __syncwarp();
      for (int i = lane; i < 512; i += 32) ovl_g[i] = P[i];
      adj = 2;
    }

    if (lane == 0) {
      p.wstate[2 * u] = (uint8_t)win_shape;
      p.wstate[2 * u + 1] = (uint8_t)win_seq;
      p.qshift_adj[u] = (int8_t)adj;
    }
    __syncwarp();
  }
}

size_t imdct_smem_bytes() { return sizeof(BlockSmem); }

cudaError_t launch_imdct(const ImdctArgs &args, int num_sms, cudaStream_t stream) {
  static xb::PerDeviceOnce configured;
  size_t smem = sizeof(BlockSmem);
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(imdct_ola_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.done();
  }
  int blocks_per_sm = (int)((227 * 1024) / (smem + 1024));
  if (blocks_per_sm < 1) blocks_per_sm = 1;
  long long need = (args.n_units + kWarpsPerBlock - 1) / kWarpsPerBlock;
  long long grid = (long long)num_sms * blocks_per_sm;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  imdct_ola_kernel<<<(unsigned)grid, kWarpsPerBlock * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace xb
