// esbr_synth_kernel.cu — the 64-band eSBR QMF synthesis bank for sm_100a (B200).
//
// One warp owns one unit (one output channel of one frame: 32 time slots x 64 complex float bands -> 2048 float samples).
// Replaces, bit-exactly, the per-slot core of
//   ixheaacd_esbr_synthesis_filt_block        decoder/ixheaacd_sbr_dec.c:583-654   (stereo_config_idx <= 0, 64 channels)
// i.e. float -> WORD32 (x 64), ixheaacd_esbr_inv_modulation (decoder/ixheaacd_qmf_dec.c:733) = ixheaacd_esbr_cos_sin_mod
// with ixheaacd_esbr_radix4bfly / ixheaacd_esbr_postradixcompute2 (decoder/generic/ixheaacd_qmf_dec_generic.c:1163-1461,
// 880-973, 975-1057), ixheaacd_shiftrountine_with_rnd_hq (:1704), ixheaacd_esbr_qmfsyn64_winadd (:1544), WORD32 -> float.
// The arithmetic between the two conversions is integer (WORD32 data and twiddles, WORD64 products), so the float
// output is bit-identical too.
//
// lane = TIME SLOT for the modulation: the 32 slots of a unit are 32 independent 2 x 32-point transforms, each run
// serially by one lane in place in its own 128-word row of shared memory (row stride 129 words: conflict-free); the
// digit-reversed radix-2 pass goes through registers.  The slot's row then receives its 128 WORD32 state samples, so
// the rows double as the window history; the 9 older blocks of the state ring stay in registers (a lane only needs its own taps).
// lane = output sample pair for the 10-tap window, in the time-invariant form the reference's ring / coefficient
// bookkeeping collapses to when both are in lock step (see sbr_lp_kernel.cu); other states take the literal form.
// Algorithmic HBM bytes per unit: 16 384 (float matrix) + 5120 + 5120 (WORD32 state in / out) + 8192 (float out) = 34 816.
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"

namespace xb {

constexpr int kEsWarps = 13;
constexpr int ES = 129;  // row stride in words

struct EsTab {
  i32 qmf_c[1280];
  i32 w32[60];
  i32 sincos[64];
  i32 alt[32];
  i32 w16[24];
  i32 sincos32[32];
  i32 alt32[16];
  i32 tcos32[64];
};
struct EsWarpS {
  i32 rows[32 * ES];  // row s: slot s (matrix row, then its 128 state samples); the 9 older state blocks stay in registers
};
struct EsBlockS {
  EsTab tab;
  EsWarpS w[kEsWarps];
};

XB_DEV i32 padd(i32 a, i32 w1, i32 b, i32 w2) { return (i32)(((long long)a * w1 + (long long)b * w2) >> 32); }
XB_DEV i32 psubw(i32 a, i32 w1, i32 b, i32 w2) {
  return (i32)((long long)((unsigned long long)((long long)a * w1) - (unsigned long long)((long long)b * w2)) >> 32);
}
XB_DEV i32 psub(i32 a, i32 w1, i32 b, i32 w2) {  // ixheaac_sub64_sat(a w1, b w2) >> 32
  const long long x = (long long)a * w1, y = (long long)b * w2;
  long long d = (long long)((unsigned long long)x - (unsigned long long)y);
  if (((x ^ y) & (x ^ d)) < 0) d = x < 0 ? (long long)0x8000000000000000ULL : 0x7fffffffffffffffLL;
  return (i32)(d >> 32);
}

// SAT = false: the unit's magnitude bound proves that no add / sub / negate of the transform can saturate (see the kernel),
// so they run as plain wrapping instructions (1 SASS instruction instead of 5) with identical results
template <bool SAT> XB_DEV i32 tadd(i32 a, i32 b) { return SAT ? add_sat(a, b) : wadd(a, b); }
template <bool SAT> XB_DEV i32 tsub(i32 a, i32 b) { return SAT ? sub_sat(a, b) : wsub(a, b); }
template <bool SAT> XB_DEV i32 tneg(i32 a) { return SAT ? neg_sat(a) : wneg(a); }
template <bool SAT> XB_DEV i32 tpsub(i32 a, i32 w1, i32 b, i32 w2) { return SAT ? psub(a, w1, b, w2) : psubw(a, w1, b, w2); }

// generic:880-973, in place on interleaved complex x (one lane)
template <bool SAT>
XB_DEV void es_radix4(const i32 *w, i32 *x, int groups, int span) {
#pragma unroll 1
  for (int g = 0; g < groups; g++) {
#pragma unroll 1
    for (int i = 0; i < span; i++) {
      i32 *e0 = x + 2 * (g * 4 * span + i), *e1 = e0 + 2 * span, *e2 = e0 + 4 * span, *e3 = e0 + 6 * span;
      const i32 *tw = w + 6 * i;
      const i32 si1 = tw[0], co1 = tw[1], si2 = tw[2], co2 = tw[3], si3 = tw[4], co3 = tw[5];
      const i32 a0 = e0[0], a1 = e0[1], b0 = e1[0], b1 = e1[1], c0 = e2[0], c1 = e2[1], d0 = e3[0], d1 = e3[1];
      const i32 xh0 = tadd<SAT>(a0, c0), xl0 = tsub<SAT>(a0, c0), xh20 = tadd<SAT>(b0, d0), xl20 = tsub<SAT>(b0, d0);
      const i32 xh1 = tadd<SAT>(a1, c1), xl1 = tsub<SAT>(a1, c1), xh21 = tadd<SAT>(b1, d1), xl21 = tsub<SAT>(b1, d1);
      const i32 xt0 = tsub<SAT>(xh0, xh20), yt0 = tsub<SAT>(xh1, xh21);
      const i32 xt1 = tadd<SAT>(xl0, xl21), xt2 = tsub<SAT>(xl0, xl21);
      const i32 yt2 = tadd<SAT>(xl1, xl20), yt1 = tsub<SAT>(xl1, xl20);
      e0[0] = tadd<SAT>(xh0, xh20);
      e0[1] = tadd<SAT>(xh1, xh21);
      e3[0] = lsl(padd(yt2, si3, xt2, co3), 1);
      e3[1] = lsl(psubw(yt2, co3, xt2, si3), 1);
      e2[0] = lsl(padd(yt0, si2, xt0, co2), 1);
      e2[1] = lsl(psubw(yt0, co2, xt0, si2), 1);
      e1[0] = lsl(padd(yt1, si1, xt1, co1), 1);
      e1[1] = lsl(psubw(yt1, co1, xt1, si1), 1);
    }
  }
}

// ixheaacd_esbr_cos_sin_mod for 64 channels on one slot row sb[0..63] | sb[64..127], in place (one lane)
template <bool SAT>
XB_DEV void es_cos_sin_mod(const EsTab &t, i32 *sb) {
  i32 *s1 = sb, *s2 = sb + 64;
  // pre-twiddle (generic:1200-1295), two steps at a time: steps n (even) and n + 1 read and write the same four words
#pragma unroll 1
  for (int n = 0; n < 32; n += 2) {
    const i32 wim0 = t.sincos[2 * n], wre0 = t.sincos[2 * n + 1], wim1 = t.sincos[2 * n + 2], wre1 = t.sincos[2 * n + 3];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      i32 *s = h ? s2 : s1;
      const i32 a = s[n], b = s[63 - n], a1 = s[n + 1], b1 = s[62 - n];
      if (!h) {
        s[n] = padd(a, wre0, b, wim0);
        s[n + 1] = tpsub<SAT>(b, wre0, a, wim0);
        s[63 - n] = tpsub<SAT>(a1, wre1, b1, wim1);
        s[62 - n] = padd(b1, wre1, a1, wim1);
      } else {
        s[n] = tpsub<SAT>(b, wim0, a, wre0);
        s[n + 1] = padd(a, wim0, b, wre0);
        s[63 - n] = padd(b1, wim1, a1, wre1);
        s[62 - n] = tpsub<SAT>(a1, wim1, b1, wre1);
      }
    }
  }
#pragma unroll 1
  for (int h = 0; h < 2; h++) {
    i32 *x = sb + 64 * h;
    es_radix4<SAT>(t.w32, x, 1, 8);
    es_radix4<SAT>(t.w32 + 48, x, 4, 2);
    // generic:975-1057 — final radix-2 with digit-reversed scatter, through registers (dig_rev_table2_32 = {0,64,16,80})
    i32 v[64];
#pragma unroll
    for (int i = 0; i < 64; i++) v[i] = x[i];
#pragma unroll
    for (int blk = 0; blk < 4; blk++) {
      const int h2 = (blk & 1) * 16 + (blk >> 1) * 4;
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const int c = (blk >> 1) * 32 + (blk & 1) * 8 + 16 * half;
        const int q = h2 + 2 * half;
        x[q] = tadd<SAT>(v[c], v[c + 2]);
        x[q + 1] = tadd<SAT>(v[c + 1], v[c + 3]);
        x[32 + q] = tsub<SAT>(v[c], v[c + 2]);
        x[32 + q + 1] = tsub<SAT>(v[c + 1], v[c + 3]);
        x[8 + q] = tadd<SAT>(v[c + 4], v[c + 6]);
        x[8 + q + 1] = tadd<SAT>(v[c + 5], v[c + 7]);
        x[40 + q] = tsub<SAT>(v[c + 4], v[c + 6]);
        x[40 + q + 1] = tsub<SAT>(v[c + 5], v[c + 7]);
      }
    }
  }
  // post-twiddle (generic:1365-1460) in place; the back pair of the next step is fetched before this step overwrites it
  {
    const i32 f10 = s1[0], f11 = s1[1], f20 = s2[0], f21 = s2[1];
    i32 re1 = s1[63], im1 = s1[62], re2 = s2[63], im2 = s2[62];
    s1[0] = f10 >> 1;
    s1[63] = tneg<SAT>(f11 >> 1);
    s2[63] = tneg<SAT>(f20 >> 1);
    s2[0] = f21 >> 1;
#pragma unroll 1
    for (int u = 0; u < 16; u++) {
      const i32 wim = t.alt[2 * u], wre = t.alt[2 * u + 1];
      i32 nre1 = 0, nim1 = 0, nre2 = 0, nim2 = 0;
      if (u + 1 < 16) {
        nre1 = s1[61 - 2 * u]; nim1 = s1[60 - 2 * u];
        nre2 = s2[61 - 2 * u]; nim2 = s2[60 - 2 * u];
      }
      s1[62 - 2 * u] = padd(re1, wre, im1, wim);
      s1[1 + 2 * u] = tpsub<SAT>(im1, wre, re1, wim);
      s2[1 + 2 * u] = tneg<SAT>(padd(re2, wre, im2, wim));
      s2[62 - 2 * u] = tpsub<SAT>(re2, wim, im2, wre);
      if (u + 1 < 16) {
        i32 fim = s1[2 + 2 * u], fre = s1[3 + 2 * u];
        s1[2 + 2 * u] = padd(fre, wim, fim, wre);
        s1[61 - 2 * u] = tpsub<SAT>(fim, wim, fre, wre);
        fim = s2[2 + 2 * u];
        fre = s2[3 + 2 * u];
        s2[61 - 2 * u] = tneg<SAT>(padd(fre, wim, fim, wre));
        s2[2 + 2 * u] = tpsub<SAT>(fre, wre, fim, wim);
      }
      re1 = nre1; im1 = nim1; re2 = nre2; im2 = nim2;
    }
  }
  // ixheaacd_shiftrountine_with_rnd_hq (generic:1704-1734), len = 64, shift = 6: row -> 128 state samples, in place
#pragma unroll 1
  for (int j = 0; j < 32; j++) {
    const i32 r1 = sb[j], i1 = sb[64 + j], r2 = sb[63 - j], i2 = sb[127 - j];
    sb[127 - j] = shl32_sat(tadd<SAT>(i1, r1), 6);
    sb[63 - j] = shl32_sat(tsub<SAT>(i2, r2), 6);
    sb[j] = shl32_sat(tsub<SAT>(i1, r1), 6);
    sb[64 + j] = shl32_sat(tadd<SAT>(i2, r2), 6);
  }
}

XB_DEV i32 f2i_x86(float v) {  // (WORD32)v as x86 CVTTSS2SI does it: out-of-range and NaN give INT_MIN
  return (fabsf(v) < 2147483648.0f) ? __float2int_rz(v) : (i32)0x80000000;
}

__global__ void __launch_bounds__(kEsWarps * 32, 1) esbr_synth_kernel(EsbrSynthArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EsBlockS &sm = *reinterpret_cast<EsBlockS *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const i32 *src = reinterpret_cast<const i32 *>(p.rom);
    i32 *dst = reinterpret_cast<i32 *>(&sm.tab);
    for (int i = threadIdx.x; i < (int)(sizeof(EsTab) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const EsTab &tab = sm.tab;
  i32 *rows = sm.w[warp].rows;
  const long long warps_total = (long long)gridDim.x * kEsWarps;
  for (long long u = (long long)blockIdx.x * kEsWarps + warp; u < p.n_units; u += warps_total) {
    __syncwarp();
    const int off0 = p.pos[2 * u], fpos0 = p.pos[2 * u + 1];
    if ((off0 & 127) != 0 || (fpos0 & 63) != 0 || off0 < 0 || off0 >= 1280 || fpos0 < 0 || fpos0 >= 640) {
      if (lane == 0 && p.err) p.err[u] = (i32)0x80000000;  // positions the reference can never produce
      continue;
    }
    const int b0 = off0 >> 7;
    {  // pull this warp's next unit towards L2 while this one is processed
      const long long un = u + warps_total;
      if (un < p.n_units) {
        if (p.qmf) {
          const char *q0 = reinterpret_cast<const char *>(p.qmf + un * 4096);
          for (int o = lane * 128; o < 16384; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q0 + o));
        }
        const char *q1 = reinterpret_cast<const char *>(p.states + un * 1280);
        for (int o = lane * 128; o < 5120; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q1 + o));
      }
    }
    // matrix rows: float -> WORD32 (sbr_dec.c:584-587), coalesced 512-byte row loads
    u32 mx = 0;  // OR of |v| (|v| - 1 for negative v) over the unit: the magnitude bound of the transform input
    if (!p.rg_par) {
      const float4 *src = reinterpret_cast<const float4 *>(p.qmf + u * 4096);
#pragma unroll 4
      for (int s = 0; s < 32; s++) {
        const float4 v = __ldg(src + 32 * s + lane);
        i32 *r = rows + ES * s + 4 * lane;
        r[0] = f2i_x86(__fmul_rn(v.x, 64.f));
        r[1] = f2i_x86(__fmul_rn(v.y, 64.f));
        r[2] = f2i_x86(__fmul_rn(v.z, 64.f));
        r[3] = f2i_x86(__fmul_rn(v.w, 64.f));
        mx |= (u32)(r[0] ^ (r[0] >> 31)) | (u32)(r[1] ^ (r[1] >> 31)) | (u32)(r[2] ^ (r[2] >> 31)) | (u32)(r[3] ^ (r[3] >> 31));
      }
    } else {  // stage mode: ixheaacd_esbr_synthesis_regrp in the load — low band from qmf_buf, high band from sbr_qmf_out
      const int xo_first = p.rg_par[4 * u], xo_rest = p.rg_par[4 * u + 1], stop = p.rg_par[4 * u + 2];
      const float *lo = (lane < 16 ? p.rg_low_re : p.rg_low_im) + u * p.rg_low_stride + 128 + 4 * (lane & 15);
      const float *hi = (lane < 16 ? p.rg_high_re : p.rg_high_im) + u * 2560 + 128 + 4 * (lane & 15);
      const int k0 = 4 * (lane & 15);
#pragma unroll 4
      for (int s = 0; s < 32; s++) {
        const int xo = s < stop ? xo_first : xo_rest;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (k0 < xo) a = *reinterpret_cast<const float4 *>(lo + 64 * s);
        if (k0 + 3 >= xo) b = *reinterpret_cast<const float4 *>(hi + 64 * s);
        i32 *r = rows + ES * s + 4 * lane;
        r[0] = f2i_x86(__fmul_rn(k0 < xo ? a.x : b.x, 64.f));
        r[1] = f2i_x86(__fmul_rn(k0 + 1 < xo ? a.y : b.y, 64.f));
        r[2] = f2i_x86(__fmul_rn(k0 + 2 < xo ? a.z : b.z, 64.f));
        r[3] = f2i_x86(__fmul_rn(k0 + 3 < xo ? a.w : b.w, 64.f));
        mx |= (u32)(r[0] ^ (r[0] >> 31)) | (u32)(r[1] ^ (r[1] >> 31)) | (u32)(r[2] ^ (r[2] >> 31)) | (u32)(r[3] ^ (r[3] >> 31));
      }
    }
    // old ring blocks: the block of age a0 (1..9) is ring block (b0 + a0) mod 10; a lane only ever needs words lane,
    // lane + 32, lane + 64, lane + 96 of each (its own window taps), so the history lives in registers
    const i32 *ss = p.states + u * 1280;
    i32 hr[9][4];
#pragma unroll
    for (int a0 = 1; a0 <= 9; a0++) {
      int b = b0 + a0;
      if (b >= 10) b -= 10;
#pragma unroll
      for (int j = 0; j < 4; j++) hr[a0 - 1][j] = __ldg(ss + 128 * b + 32 * j + lane);
    }
    __syncwarp();
    // The transform grows magnitudes by at most 8 x 8 x 2 = 128 (two radix-4 stages whose twiddled outputs are
    // lsl((a w1 + b w2) >> 32, 1) <= 8 x input, one radix-2 stage; pre- and post-twiddles do not grow) and the state
    // conversion adds two outputs: with every |v| <= 2^22 nothing reaches 2^31, no saturating operation can saturate and
    // the wrapping variant computes the same words.
    if (__reduce_or_sync(0xffffffffu, mx) < (1u << 22))
      es_cos_sin_mod<false>(tab, rows + ES * lane);
    else
      es_cos_sin_mod<true>(tab, rows + ES * lane);
    __syncwarp();
    // 10-tap window (generic:1544-1575): lane -> outputs lane and lane + 32 of every slot (stride-1 rows: conflict-free)
    float *out = p.out ? p.out + u * 2048 : nullptr;
    int16_t *pcm = p.pcm16 ? p.pcm16 + ((u / p.pcm_ch_fac) * 2048) * p.pcm_ch_fac + (u % p.pcm_ch_fac) : nullptr;
    const int pstep = p.pcm_ch_fac;
    auto emit = [&](int idx, i32 o) {  // WORD32 -> float (sbr_dec.c:650); optional ixheaacd_samples_sat (decode_main.c:82-104)
      const float f = __fmul_rn(__int2float_rn(o), 1.0f / 65536.0f);
      if (out) out[idx] = f;
      if (pcm) {
        const float c = f > 32767.0f ? 32767.0f : (f < -32768.0f ? -32768.0f : f);
        pcm[idx * pstep] = (int16_t)__float2int_rz(c);
      }
    };
    const int f0s = fpos0 >> 6;
    const bool lock = p.periodic && (off0 & 255) == 0 && (fpos0 & 127) == 0 && ((b0 + f0s) % 10 == 0);
    if (lock) {
      i32 c0[10], c1[10];
#pragma unroll
      for (int a = 0; a < 10; a++) {
        c0[a] = tab.qmf_c[64 * a + lane];
        c1[a] = tab.qmf_c[64 * a + 32 + lane];
      }
      const i32 *hp = rows + lane;
#pragma unroll
      for (int i = 0; i < 9; i++) {  // slots whose window still reaches into the old state (registers)
        unsigned long long acc0 = 0, acc1 = 0;
#pragma unroll
        for (int a = 0; a < 10; a++) {
          i32 q0, q1;
          if (a <= i) {
            const i32 *q = hp + ES * (i - a) + 64 * (a & 1);
            q0 = q[0];
            q1 = q[32];
          } else {
            q0 = hr[a - i - 1][2 * (a & 1)];
            q1 = hr[a - i - 1][2 * (a & 1) + 1];
          }
          acc0 += (unsigned long long)((long long)q0 * c0[a]);
          acc1 += (unsigned long long)((long long)q1 * c1[a]);
        }
        const i32 o0 = (i32)((long long)acc0 >> 31), o1 = (i32)((long long)acc1 >> 31);
        emit(64 * i + lane, o0);
        emit(64 * i + 32 + lane, o1);
      }
#pragma unroll 1
      for (int i = 9; i < 32; i++) {
        unsigned long long acc0 = 0, acc1 = 0;
#pragma unroll
        for (int a = 0; a < 10; a++) {
          const i32 *q = hp + ES * (i - a) + 64 * (a & 1);
          acc0 += (unsigned long long)((long long)q[0] * c0[a]);
          acc1 += (unsigned long long)((long long)q[32] * c1[a]);
        }
        const i32 o0 = (i32)((long long)acc0 >> 31), o1 = (i32)((long long)acc1 >> 31);
        emit(64 * i + lane, o0);
        emit(64 * i + 32 + lane, o1);
      }
    } else {
      int fpos = fpos0;
#pragma unroll 1
      for (int i = 0; i < 32; i++) {
        unsigned long long acc0 = 0, acc1 = 0;
        int ab = (i - b0) % 10;
        if (ab < 0) ab += 10;
#pragma unroll 1
        for (int b = 0; b < 10; b++) {
          int a = ab + b;
          if (a >= 10) a -= 10;
          int hb = b0 + a - i;  // ring block of the old state when the tap is older than this frame
          if (hb >= 10) hb -= 10;
          const i32 *q = (i >= a ? rows + ES * (i - a) : ss + 128 * hb) + 64 * ((i + b) & 1) + lane;
          const i32 *c = tab.qmf_c + fpos + 64 * b + lane;
          acc0 += (unsigned long long)((long long)q[0] * c[0]);
          acc1 += (unsigned long long)((long long)q[32] * c[32]);
        }
        const i32 o0 = (i32)((long long)acc0 >> 31), o1 = (i32)((long long)acc1 >> 31);
        emit(64 * i + lane, o0);
        emit(64 * i + 32 + lane, o1);
        fpos += 64;
        if (fpos == 640) fpos = 0;
      }
    }
    // ring after 32 slots: block b holds slot 31 - a, a = (b - b0 + 31) mod 10
    {
      i32 *sd = p.states + u * 1280;
#pragma unroll 1
      for (int i = lane; i < 1280; i += 32) {
        const int b = i >> 7;
        int a = (b - b0 + 31) % 10;
        if (a < 0) a += 10;
        sd[i] = rows[ES * (31 - a) + (i & 127)];
      }
      if (lane == 0) {
        int off = (off0 - 128 * 32) % 1280;
        if (off < 0) off += 1280;
        p.pos[2 * u] = off;
        p.pos[2 * u + 1] = (fpos0 + 64 * 32) % 640;
        if (p.err) p.err[u] = 0;
      }
    }
  }
}


// =====================================================================================================================
// eSBR 32-band analysis bank: ixheaacd_esbr_analysis_filt_block (decoder/ixheaacd_sbr_dec.c:185-295) for 32 channels and
// 32 time slots: float -> WORD32 (x 2^15), ixheaacd_esbr_qmfanal32_winadd (decoder/ixheaacd_qmf_dec.c:537-640),
// ixheaacd_esbr_fwd_modulation (generic:1463-1506) = >> 4, fold, ixheaacd_esbr_cos_sin_mod (M = 16, esbr_w_16,
// ixheaacd_esbr_postradixcompute4), t_cos rotation, WORD32 -> float (x 1/256).
// Window: lane = output, the time-invariant form of the ring bookkeeping (lock step) on a flat WORD32 history; modulation:
// lane = slot, in place in the slot's row; rotation + conversion + coalesced store: lane = band.
// Algorithmic HBM bytes per unit: 4096 (float in) + 1280 + 1280 (WORD32 ring in / out) + 8192 (32 x 32 complex float out).
// =====================================================================================================================
constexpr int kEaWarps = 12;
constexpr int EA = 97;  // row stride: 96 words used (s1 at 0..31, window output at 0..63, s2 at 64..95)
struct EaWarpS {
  i32 T[1312];       // T[j] = sample at time j - 288 relative to the frame start (lock-step path); ring in the literal path
  i32 rows[32 * EA];
};
struct EaBlockS {
  EsTab tab;
  EaWarpS w[kEaWarps];
};

// ixheaacd_esbr_cos_sin_mod for 32 channels (M = 16) on one slot row: s1 = sb[0..31], s2 = sb[64..95], in place
template <bool SAT>
XB_DEV void ea_cos_sin_mod(const EsTab &t, i32 *sb) {
  i32 *s1 = sb, *s2 = sb + 64;
#pragma unroll 1
  for (int n = 0; n < 16; n += 2) {
    const i32 wim0 = t.sincos32[2 * n], wre0 = t.sincos32[2 * n + 1], wim1 = t.sincos32[2 * n + 2], wre1 = t.sincos32[2 * n + 3];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      i32 *s = h ? s2 : s1;
      const i32 a = s[n], b = s[31 - n], a1 = s[n + 1], b1 = s[30 - n];
      if (!h) {
        s[n] = padd(a, wre0, b, wim0);
        s[n + 1] = tpsub<SAT>(b, wre0, a, wim0);
        s[31 - n] = tpsub<SAT>(a1, wre1, b1, wim1);
        s[30 - n] = padd(b1, wre1, a1, wim1);
      } else {
        s[n] = tpsub<SAT>(b, wim0, a, wre0);
        s[n + 1] = padd(a, wim0, b, wre0);
        s[31 - n] = padd(b1, wim1, a1, wre1);
        s[30 - n] = tpsub<SAT>(a1, wim1, b1, wre1);
      }
    }
  }
#pragma unroll 1
  for (int h = 0; h < 2; h++) {
    i32 *x = sb + 64 * h;
    es_radix4<SAT>(t.w16, x, 1, 4);
    // generic:1059-1161 — final radix-4 (no twiddles) with digit-reversed scatter (dig_rev_table4_16 = {0, 16}), via registers
    i32 v[32];
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = x[i];
#pragma unroll
    for (int k = 0; k < 2; k++) {
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const int c = 16 * k + 8 * half, o = 4 * k + 2 * half;
        const i32 xh0 = tadd<SAT>(v[c], v[c + 4]), xh1 = tadd<SAT>(v[c + 1], v[c + 5]);
        const i32 xl0 = tsub<SAT>(v[c], v[c + 4]), xl1 = tsub<SAT>(v[c + 1], v[c + 5]);
        const i32 zh0 = tadd<SAT>(v[c + 2], v[c + 6]), zh1 = tadd<SAT>(v[c + 3], v[c + 7]);
        const i32 zl0 = tsub<SAT>(v[c + 2], v[c + 6]), zl1 = tsub<SAT>(v[c + 3], v[c + 7]);
        x[o] = tadd<SAT>(xh0, zh0);
        x[o + 1] = tadd<SAT>(xh1, zh1);
        x[8 + o] = tadd<SAT>(xl0, zl1);
        x[8 + o + 1] = tsub<SAT>(xl1, zl0);
        x[16 + o] = tsub<SAT>(xh0, zh0);
        x[16 + o + 1] = tsub<SAT>(xh1, zh1);
        x[24 + o] = tsub<SAT>(xl0, zl1);
        x[24 + o + 1] = tadd<SAT>(xl1, zl0);
      }
    }
  }
  {  // post-twiddle (generic:1365-1460), N = 32, H = 8, in place
    const i32 f10 = s1[0], f11 = s1[1], f20 = s2[0], f21 = s2[1];
    i32 re1 = s1[31], im1 = s1[30], re2 = s2[31], im2 = s2[30];
    s1[0] = f10 >> 1;
    s1[31] = tneg<SAT>(f11 >> 1);
    s2[31] = tneg<SAT>(f20 >> 1);
    s2[0] = f21 >> 1;
#pragma unroll 1
    for (int u = 0; u < 8; u++) {
      const i32 wim = t.alt32[2 * u], wre = t.alt32[2 * u + 1];
      i32 nre1 = 0, nim1 = 0, nre2 = 0, nim2 = 0;
      if (u + 1 < 8) {
        nre1 = s1[29 - 2 * u]; nim1 = s1[28 - 2 * u];
        nre2 = s2[29 - 2 * u]; nim2 = s2[28 - 2 * u];
      }
      s1[30 - 2 * u] = padd(re1, wre, im1, wim);
      s1[1 + 2 * u] = tpsub<SAT>(im1, wre, re1, wim);
      s2[1 + 2 * u] = tneg<SAT>(padd(re2, wre, im2, wim));
      s2[30 - 2 * u] = tpsub<SAT>(re2, wim, im2, wre);
      if (u + 1 < 8) {
        i32 fim = s1[2 + 2 * u], fre = s1[3 + 2 * u];
        s1[2 + 2 * u] = padd(fre, wim, fim, wre);
        s1[29 - 2 * u] = tpsub<SAT>(fim, wim, fre, wre);
        fim = s2[2 + 2 * u];
        fre = s2[3 + 2 * u];
        s2[29 - 2 * u] = tneg<SAT>(padd(fre, wim, fim, wre));
        s2[2 + 2 * u] = tpsub<SAT>(fre, wre, fim, wim);
      }
      re1 = nre1; im1 = nim1; re2 = nre2; im2 = nim2;
    }
  }
}

__global__ void __launch_bounds__(kEaWarps * 32, 1) esbr_anal_kernel(EsbrAnalArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EaBlockS &sm = *reinterpret_cast<EaBlockS *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const i32 *src = reinterpret_cast<const i32 *>(p.rom);
    i32 *dst = reinterpret_cast<i32 *>(&sm.tab);
    for (int i = threadIdx.x; i < (int)(sizeof(EsTab) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const EsTab &tab = sm.tab;
  EaWarpS &w = sm.w[warp];
  const long long warps_total = (long long)gridDim.x * kEaWarps;
  const i32 tc = tab.tcos32[2 * lane], ts = tab.tcos32[2 * lane + 1];
  for (long long u = (long long)blockIdx.x * kEaWarps + warp; u < p.n_units; u += warps_total) {
    __syncwarp();
    const float *tin = p.time_in ? p.time_in + u * 1024 : nullptr;
    const i32 *cin = p.core_in ? p.core_in + u * 1024 : nullptr;
    const int16_t *pin = p.pcm_in ? p.pcm_in + ((u / p.pcm_ch_fac) * 1024) * p.pcm_ch_fac + (u % p.pcm_ch_fac) : nullptr;
    const int pstep = p.pcm_ch_fac;
    auto sample = [&](int j) -> float {  // the core-coder sample as the float the reference hands to the bank
      if (tin) return __ldg(tin + j);
      if (cin) return __fmul_rn(__int2float_rn(__ldg(cin + j)), 0.000030517578125f);
      return __int2float_rn((int)__ldg(pin + j * pstep));
    };
    i32 *ring = p.states + u * 320;
    int pos = p.pos[2 * u], f1 = p.pos[2 * u + 1], f2 = f1 + 64;
    if ((pos & 31) != 0 || pos < 0 || pos > 288 || (f1 & 63) != 0 || f1 < 0 || f1 > 576) {
      if (lane == 0 && p.err) p.err[u] = (i32)0x80000000;
      continue;
    }
    if (u + warps_total < p.n_units) {  // pull the next unit's input and ring into L2 while this one computes
      if (tin || cin) {
        const char *q0 = tin ? reinterpret_cast<const char *>(tin + warps_total * 1024)
                             : reinterpret_cast<const char *>(cin + warps_total * 1024);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q0 + lane * 128));
      }
      const char *q1 = reinterpret_cast<const char *>(ring + warps_total * 320);
      if (lane < 10) asm volatile("prefetch.global.L2 [%0];" ::"l"(q1 + lane * 128));
    }
    const int P0 = pos >> 5, F0 = f1 >> 6;
    const bool lock = p.periodic && (pos & 63) == 0 && (f1 & 127) == 0 && pos <= 256 && f1 <= 512 && ((P0 + F0) % 10 == 0);
    if (lock) {
      // flat history: 288 old samples from the ring (block q holds slot P0 - q mod 10, newest sample first), 1024 new ones
      {
        i32 vh[9];  // the nine old blocks: all requests in flight
#pragma unroll
        for (int r = 0; r < 9; r++) {
          int q = P0 + 9 - r;
          if (q >= 10) q -= 10;
          vh[r] = ring[32 * q + 31 - lane];
        }
#pragma unroll
        for (int r = 0; r < 9; r++) w.T[32 * r + lane] = vh[r];
      }
      {
        float v[32];  // all 32 loads in flight before the first conversion
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = sample(32 * j + lane);
#pragma unroll
        for (int j = 0; j < 32; j++) w.T[288 + 32 * j + lane] = f2i_x86(__fmul_rn(v[j], 32768.0f));
      }
      i32 ca[5], cb[5];
#pragma unroll
      for (int m = 0; m < 5; m++) {
        ca[m] = tab.qmf_c[128 * m + 2 * lane];
        cb[m] = tab.qmf_c[64 + 128 * m + 2 * lane];
      }
      __syncwarp();
      const i32 *ta = w.T + 288 + 31 - lane, *tb = w.T + 288 - 1 - lane;
#pragma unroll 2
      for (int slot = 0; slot < 32; slot++) {
        unsigned long long a = 0, b = 0;
#pragma unroll
        for (int m = 0; m < 5; m++) {
          a += (unsigned long long)((long long)ta[32 * slot - 64 * m] * ca[m]);
          b += (unsigned long long)((long long)tb[32 * slot - 64 * m] * cb[m]);
        }
        w.rows[EA * slot + lane] = (i32)((long long)a >> 31);
        w.rows[EA * slot + 32 + lane] = (i32)((long long)b >> 31);
      }
      // ring after 32 slots
#pragma unroll 1
      for (int pp = lane; pp < 320; pp += 32) {
        const int q = pp >> 5;
        int d = (31 - (P0 - q)) % 10;
        if (d < 0) d += 10;
        ring[pp] = w.T[288 + 32 * (31 - d) + 31 - (pp & 31)];
      }
      const int Pf = (P0 + 8) % 10;
      pos = 32 * Pf;
      f1 = 64 * ((10 - Pf) % 10);
    } else {
      // literal ring emulation (sbr_dec.c:244-276) for (position, phase) pairs the reference never produces itself
      i32 *rg = w.T;
      for (int j = lane; j < 320; j += 32) rg[j] = ring[j];
      __syncwarp();
#pragma unroll 1
      for (int slot = 0; slot < 32; slot++) {
        rg[pos + 31 - lane] = f2i_x86(__fmul_rn(sample(32 * slot + lane), 32768.0f));
        __syncwarp();
        const i32 *fp1 = rg + ((slot & 1) ? 32 : 0), *fp2 = rg + ((slot & 1) ? 0 : 32);
        unsigned long long a = 0, b = 0;
#pragma unroll
        for (int j = 0; j < 5; j++) {
          a += (unsigned long long)((long long)fp1[lane + 64 * j] * tab.qmf_c[f1 + 2 * (lane + 64 * j)]);
          b += (unsigned long long)((long long)fp2[lane + 64 * j] * tab.qmf_c[f2 + 2 * (lane + 64 * j)]);
        }
        __syncwarp();
        pos -= 32;
        if (pos < 0) pos = 288;
        {
          const int n1 = f2 + 64, n2 = f1 + 64;
          f1 = n1;
          f2 = n2;
          if (f2 > 640) {
            f1 = 0;
            f2 = 64;
          }
        }
        w.rows[EA * slot + lane] = (i32)((long long)a >> 31);
        w.rows[EA * slot + 32 + lane] = (i32)((long long)b >> 31);
      }
      for (int j = lane; j < 320; j += 32) ring[j] = rg[j];
    }
    __syncwarp();
    {  // lane = slot: fold (generic:1475-1482) in place, then the modulation
      i32 *sb = w.rows + EA * lane;
      u32 mx = 0;
#pragma unroll 1
      for (int k = 0; k < 16; k++) {
        const i32 a0 = sb[k] >> 4, a1 = sb[63 - k] >> 4, b0 = sb[31 - k] >> 4, b1 = sb[32 + k] >> 4;
        const i32 v0 = sub_sat(a0, a1), v1 = add_sat(a0, a1), v2 = sub_sat(b0, b1), v3 = add_sat(b0, b1);
        sb[k] = v0;
        sb[64 + k] = v1;
        sb[31 - k] = v2;
        sb[64 + 31 - k] = v3;
        mx |= (u32)(v0 ^ (v0 >> 31)) | (u32)(v1 ^ (v1 >> 31)) | (u32)(v2 ^ (v2 >> 31)) | (u32)(v3 ^ (v3 >> 31));
      }
      // growth of the 16-point transform: <= 8 (radix-4 with twiddles) x 4 (plain radix-4) = 32; below 2^25 nothing saturates
      if (__reduce_or_sync(0xffffffffu, mx) < (1u << 25))
        ea_cos_sin_mod<false>(tab, sb);
      else
        ea_cos_sin_mod<true>(tab, sb);
    }
    __syncwarp();
    // lane = band: t_cos rotation (generic:1490-1505), WORD32 -> float (x 1/256), coalesced rows of the output matrix
    if (p.stage_re) {
      const int hist = p.stage_hist_rows;
      float *sre = p.stage_re + u * (long long)(32 + hist) * 64, *sim = p.stage_im + u * (long long)(32 + hist) * 64;
      for (int r0 = 0; r0 < hist; r0 += 8) {  // history rows: 32.. -> 0.. (forward copy, 8 rows at a time), all 64 bands
        float vr[16], vi[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
          vr[q] = sre[64 * (32 + r0) + 32 * q + lane];
          vi[q] = sim[64 * (32 + r0) + 32 * q + lane];
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 16; q++) {
          sre[64 * r0 + 32 * q + lane] = vr[q];
          sim[64 * r0 + 32 * q + lane] = vi[q];
        }
        __syncwarp();
      }
#pragma unroll 2
      for (int s = 0; s < 32; s++) {
        const i32 re = w.rows[EA * s + lane], im = w.rows[EA * s + 64 + lane];
        const i32 r2 = (i32)(((long long)re * tc + (long long)im * ts) >> 31);
        const long long x = (long long)im * tc, y = (long long)re * ts;
        long long d = (long long)((unsigned long long)x - (unsigned long long)y);
        if (((x ^ y) & (x ^ d)) < 0) d = x < 0 ? (long long)0x8000000000000000ULL : 0x7fffffffffffffffLL;
        sre[64 * (hist + s) + lane] = __fmul_rn(__int2float_rn(r2), 1.0f / 256.0f);
        sim[64 * (hist + s) + lane] = __fmul_rn(__int2float_rn((i32)(d >> 31)), 1.0f / 256.0f);
      }
      if (lane == 0) {
        p.pos[2 * u] = pos;
        p.pos[2 * u + 1] = f1;
        if (p.err) p.err[u] = 0;
      }
      continue;
    }
    float *out = p.qmf + u * p.out_stride;
#pragma unroll 2
    for (int s = 0; s < 32; s++) {
      const i32 re = w.rows[EA * s + lane], im = w.rows[EA * s + 64 + lane];
      const i32 r2 = (i32)(((long long)re * tc + (long long)im * ts) >> 31);
      const long long x = (long long)im * tc, y = (long long)re * ts;
      long long d = (long long)((unsigned long long)x - (unsigned long long)y);
      if (((x ^ y) & (x ^ d)) < 0) d = x < 0 ? (long long)0x8000000000000000ULL : 0x7fffffffffffffffLL;
      out[128 * s + lane] = __fmul_rn(__int2float_rn(r2), 1.0f / 256.0f);
      out[128 * s + 64 + lane] = __fmul_rn(__int2float_rn((i32)(d >> 31)), 1.0f / 256.0f);
    }
    if (lane == 0) {
      p.pos[2 * u] = pos;
      p.pos[2 * u + 1] = f1;
      if (p.err) p.err[u] = 0;
    }
  }
}

cudaError_t launch_esbr_anal(const EsbrAnalArgs &args, int num_sms, cudaStream_t stream) {
  static xb::PerDeviceOnce configured;
  const size_t smem = sizeof(EaBlockS);
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(esbr_anal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.done();
  }
  long long need = (args.n_units + kEaWarps - 1) / kEaWarps;
  long long grid = num_sms;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  esbr_anal_kernel<<<(unsigned)grid, kEaWarps * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

size_t esbr_synth_table_bytes() { return sizeof(EsTab); }
// returns 1 if esbr_qmf_c is periodic with period 640 (the lock-step window form needs it), 0 otherwise
int esbr_synth_build_tables(const uint8_t *erom, uint8_t *out) {
  EsTab *t = reinterpret_cast<EsTab *>(out);
  const int32_t *c = reinterpret_cast<const int32_t *>(erom + kEsRomQmfC);
  memcpy(t->qmf_c, c, sizeof(t->qmf_c));
  memcpy(t->w32, erom + kEsRomW32, sizeof(t->w32));
  memcpy(t->sincos, erom + kEsRomSinCos, sizeof(t->sincos));
  memcpy(t->alt, erom + kEsRomAlt, sizeof(t->alt));
  memcpy(t->w16, erom + kEsRomW16, sizeof(t->w16));
  memcpy(t->sincos32, erom + kEsRomSinCos32, sizeof(t->sincos32));
  memcpy(t->alt32, erom + kEsRomAlt32, sizeof(t->alt32));
  memcpy(t->tcos32, erom + kEsRomTCos32, sizeof(t->tcos32));
  int periodic = 1;
  for (int i = 0; i < 640; i++)
    if (c[i] != c[i + 640]) periodic = 0;
  return periodic;
}

cudaError_t launch_esbr_synth(const EsbrSynthArgs &args, int num_sms, cudaStream_t stream) {
  static xb::PerDeviceOnce configured;
  const size_t smem = sizeof(EsBlockS);
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(esbr_synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.done();
  }
  long long need = (args.n_units + kEsWarps - 1) / kEsWarps;
  long long grid = num_sms;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  esbr_synth_kernel<<<(unsigned)grid, kEsWarps * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace xb
