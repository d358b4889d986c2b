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
// the rows double as the window history; the 9 older blocks of the state ring sit in the 9 rows before row 0.
// lane = output sample pair for the 10-tap window, in the time-invariant form the reference's ring / coefficient
// bookkeeping collapses to when both are in lock step (see sbr_lp_kernel.cu); other states take the literal form.
// Algorithmic HBM bytes per unit: 16 384 (float matrix) + 5120 + 5120 (WORD32 state in / out) + 8192 (float out) = 34 816.
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"

namespace xb {

constexpr int kEsWarps = 10;
constexpr int ES = 129;  // row stride in words

struct EsTab {
  i32 qmf_c[1280];
  i32 w32[60];
  i32 sincos[64];
  i32 alt[32];
};
struct EsWarpS {
  i32 rows[(9 + 32) * ES];  // rows 0..8: old state blocks (age 9..1), row 9 + s: slot s
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

// generic:880-973, in place on interleaved complex x (one lane)
XB_DEV void es_radix4(const i32 *w, i32 *x, int groups, int span) {
#pragma unroll 1
  for (int g = 0; g < groups; g++) {
#pragma unroll 1
    for (int i = 0; i < span; i++) {
      i32 *e0 = x + 2 * (g * 4 * span + i), *e1 = e0 + 2 * span, *e2 = e0 + 4 * span, *e3 = e0 + 6 * span;
      const i32 *tw = w + 6 * i;
      const i32 si1 = tw[0], co1 = tw[1], si2 = tw[2], co2 = tw[3], si3 = tw[4], co3 = tw[5];
      const i32 a0 = e0[0], a1 = e0[1], b0 = e1[0], b1 = e1[1], c0 = e2[0], c1 = e2[1], d0 = e3[0], d1 = e3[1];
      const i32 xh0 = add_sat(a0, c0), xl0 = sub_sat(a0, c0), xh20 = add_sat(b0, d0), xl20 = sub_sat(b0, d0);
      const i32 xh1 = add_sat(a1, c1), xl1 = sub_sat(a1, c1), xh21 = add_sat(b1, d1), xl21 = sub_sat(b1, d1);
      const i32 xt0 = sub_sat(xh0, xh20), yt0 = sub_sat(xh1, xh21);
      const i32 xt1 = add_sat(xl0, xl21), xt2 = sub_sat(xl0, xl21);
      const i32 yt2 = add_sat(xl1, xl20), yt1 = sub_sat(xl1, xl20);
      e0[0] = add_sat(xh0, xh20);
      e0[1] = add_sat(xh1, xh21);
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
        s[n + 1] = psub(b, wre0, a, wim0);
        s[63 - n] = psub(a1, wre1, b1, wim1);
        s[62 - n] = padd(b1, wre1, a1, wim1);
      } else {
        s[n] = psub(b, wim0, a, wre0);
        s[n + 1] = padd(a, wim0, b, wre0);
        s[63 - n] = padd(b1, wim1, a1, wre1);
        s[62 - n] = psub(a1, wim1, b1, wre1);
      }
    }
  }
#pragma unroll 1
  for (int h = 0; h < 2; h++) {
    i32 *x = sb + 64 * h;
    es_radix4(t.w32, x, 1, 8);
    es_radix4(t.w32 + 48, x, 4, 2);
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
        x[q] = add_sat(v[c], v[c + 2]);
        x[q + 1] = add_sat(v[c + 1], v[c + 3]);
        x[32 + q] = sub_sat(v[c], v[c + 2]);
        x[32 + q + 1] = sub_sat(v[c + 1], v[c + 3]);
        x[8 + q] = add_sat(v[c + 4], v[c + 6]);
        x[8 + q + 1] = add_sat(v[c + 5], v[c + 7]);
        x[40 + q] = sub_sat(v[c + 4], v[c + 6]);
        x[40 + q + 1] = sub_sat(v[c + 5], v[c + 7]);
      }
    }
  }
  // post-twiddle (generic:1365-1460) in place; the back pair of the next step is fetched before this step overwrites it
  {
    const i32 f10 = s1[0], f11 = s1[1], f20 = s2[0], f21 = s2[1];
    i32 re1 = s1[63], im1 = s1[62], re2 = s2[63], im2 = s2[62];
    s1[0] = f10 >> 1;
    s1[63] = neg_sat(f11 >> 1);
    s2[63] = neg_sat(f20 >> 1);
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
      s1[1 + 2 * u] = psub(im1, wre, re1, wim);
      s2[1 + 2 * u] = neg_sat(padd(re2, wre, im2, wim));
      s2[62 - 2 * u] = psub(re2, wim, im2, wre);
      if (u + 1 < 16) {
        i32 fim = s1[2 + 2 * u], fre = s1[3 + 2 * u];
        s1[2 + 2 * u] = padd(fre, wim, fim, wre);
        s1[61 - 2 * u] = psub(fim, wim, fre, wre);
        fim = s2[2 + 2 * u];
        fre = s2[3 + 2 * u];
        s2[61 - 2 * u] = neg_sat(padd(fre, wim, fim, wre));
        s2[2 + 2 * u] = psub(fre, wre, fim, wim);
      }
      re1 = nre1; im1 = nim1; re2 = nre2; im2 = nim2;
    }
  }
  // ixheaacd_shiftrountine_with_rnd_hq (generic:1704-1734), len = 64, shift = 6: row -> 128 state samples, in place
#pragma unroll 1
  for (int j = 0; j < 32; j++) {
    const i32 r1 = sb[j], i1 = sb[64 + j], r2 = sb[63 - j], i2 = sb[127 - j];
    sb[127 - j] = shl32_sat(add_sat(i1, r1), 6);
    sb[63 - j] = shl32_sat(sub_sat(i2, r2), 6);
    sb[j] = shl32_sat(sub_sat(i1, r1), 6);
    sb[64 + j] = shl32_sat(add_sat(i2, r2), 6);
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
  i32 *hist = sm.w[warp].rows;
  i32 *rows = hist + 9 * ES;
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
        const char *q0 = reinterpret_cast<const char *>(p.qmf + un * 4096);
        for (int o = lane * 128; o < 16384; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q0 + o));
        const char *q1 = reinterpret_cast<const char *>(p.states + un * 1280);
        for (int o = lane * 128; o < 5120; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q1 + o));
      }
    }
    // matrix rows: float -> WORD32 (sbr_dec.c:584-587), coalesced 512-byte row loads
    const float4 *src = reinterpret_cast<const float4 *>(p.qmf + u * 4096);
#pragma unroll 4
    for (int s = 0; s < 32; s++) {
      const float4 v = __ldg(src + 32 * s + lane);
      i32 *r = rows + ES * s + 4 * lane;
      r[0] = f2i_x86(__fmul_rn(v.x, 64.f));
      r[1] = f2i_x86(__fmul_rn(v.y, 64.f));
      r[2] = f2i_x86(__fmul_rn(v.z, 64.f));
      r[3] = f2i_x86(__fmul_rn(v.w, 64.f));
    }
    {  // old ring blocks: the block of age a0 (1..9) goes to history row 9 - a0
      const i32 *ss = p.states + u * 1280;
#pragma unroll 1
      for (int i = lane; i < 1280; i += 32) {
        const int b = i >> 7;
        int a0 = b - b0;
        if (a0 < 0) a0 += 10;
        if (a0 != 0) hist[ES * (9 - a0) + (i & 127)] = ss[i];
      }
    }
    __syncwarp();
    es_cos_sin_mod(tab, rows + ES * lane);
    __syncwarp();
    // 10-tap window (generic:1544-1575): lane -> outputs lane and lane + 32 of every slot (stride-1 rows: conflict-free)
    float *out = p.out + u * 2048;
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
#pragma unroll 1
      for (int i = 0; i < 32; i++) {
        unsigned long long acc0 = 0, acc1 = 0;
#pragma unroll
        for (int a = 0; a < 10; a++) {
          const i32 *q = hp + ES * (i - a) + 64 * (a & 1);
          acc0 += (unsigned long long)((long long)q[0] * c0[a]);
          acc1 += (unsigned long long)((long long)q[32] * c1[a]);
        }
        const i32 o0 = (i32)((long long)acc0 >> 31), o1 = (i32)((long long)acc1 >> 31);
        out[64 * i + lane] = __fmul_rn(__int2float_rn(o0), 1.0f / 65536.0f);
        out[64 * i + 32 + lane] = __fmul_rn(__int2float_rn(o1), 1.0f / 65536.0f);
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
          const i32 *q = rows + ES * (i - a) + 64 * ((i + b) & 1) + lane;
          const i32 *c = tab.qmf_c + fpos + 64 * b + lane;
          acc0 += (unsigned long long)((long long)q[0] * c[0]);
          acc1 += (unsigned long long)((long long)q[32] * c[32]);
        }
        const i32 o0 = (i32)((long long)acc0 >> 31), o1 = (i32)((long long)acc1 >> 31);
        out[64 * i + lane] = __fmul_rn(__int2float_rn(o0), 1.0f / 65536.0f);
        out[64 * i + 32 + lane] = __fmul_rn(__int2float_rn(o1), 1.0f / 65536.0f);
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

size_t esbr_synth_table_bytes() { return sizeof(EsTab); }
// returns 1 if esbr_qmf_c is periodic with period 640 (the lock-step window form needs it), 0 otherwise
int esbr_synth_build_tables(const uint8_t *erom, uint8_t *out) {
  EsTab *t = reinterpret_cast<EsTab *>(out);
  const int32_t *c = reinterpret_cast<const int32_t *>(erom + kEsRomQmfC);
  memcpy(t->qmf_c, c, sizeof(t->qmf_c));
  memcpy(t->w32, erom + kEsRomW32, sizeof(t->w32));
  memcpy(t->sincos, erom + kEsRomSinCos, sizeof(t->sincos));
  memcpy(t->alt, erom + kEsRomAlt, sizeof(t->alt));
  int periodic = 1;
  for (int i = 0; i < 640; i++)
    if (c[i] != c[i + 640]) periodic = 0;
  return periodic;
}

cudaError_t launch_esbr_synth(const EsbrSynthArgs &args, int num_sms, cudaStream_t stream) {
  static bool configured = false;
  const size_t smem = sizeof(EsBlockS);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(esbr_synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  long long need = (args.n_units + kEsWarps - 1) / kEsWarps;
  long long grid = num_sms;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  esbr_synth_kernel<<<(unsigned)grid, kEsWarps * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace xb
