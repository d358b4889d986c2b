// usac_fd_kernel.cu — USAC frequency-domain core transform (IMDCT + windowing + overlap) for sm_100a (B200).
//
// One warp owns one unit (one core channel of one USAC frame, 1024 coefficients).  Replaces, bit-exactly, for pure FD
// streams (previous frame FD, no FAC data, no concealment, ccfl = 1024):
//   ixheaacd_fd_frm_dec                 decoder/ixheaacd_imdct.c:596-654
//   ixheaacd_fd_imdct_long / _short     :477-594, :336-475
//   ixheaacd_acelp_imdct, ixheaacd_fft_based_imdct, ixheaacd_calc_pre_twid_dec, ixheaacd_calc_post_twid_dec  :186, :149, :111, :129
//   ixheaacd_complex_fft -> ixheaacd_complex_fft_p2_dec (fft_mode = 1)   decoder/ixheaacd_fft.c:2664, :1412-2491
//   ixheaacd_windowing_long1 / long3 / short2 / short3 / short4, ixheaacd_scale_down[_adj]   decoder/ixheaacd_basic_ops.c:77-657
//   ixheaacd_calc_window                decoder/ixheaacd_Windowing.c:29-111
//
// The 512-point complex FFT (or the eight 64-point FFTs of an EIGHT_SHORT frame, run as one batch) is the reference's
// own radix-4 graph — digit-reversed first stage, middle stages with four twiddle segments, a final radix-2 stage for
// 512 — because every butterfly saturates and therefore fixes the result.  128 butterflies per stage = 4 per lane.
// Spectrum and FFT workspace live in shared memory (2 x 5 KB per warp, rows padded so that the radix-4 legs of a
// half-warp fall into distinct banks); overlap and output stream through HBM once (long frames).
// Algorithmic HBM bytes per unit: 4096 (coefficients) + 4096 + 4096 (overlap in / out) + 4096 (WORD32 out) = 16 384.
#include <cstdint>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"

namespace xb {

constexpr int kUfWarps = 8;
constexpr int kUfBuf = 1280;  // 1024 words + 8 pad words per 32
constexpr int kURomPad = 15888;  // kURomBytes rounded up to 16

XB_DEV int PA(int a) { return a + ((a >> 5) << 3); }  // padded address of word a

// fft.c:48 — (a * b) >> 31 saturated: only INT_MIN * INT_MIN exceeds 32 bits (the twiddle table does contain INT_MIN)
XB_DEV i32 mul_sat31(i32 a, i32 b) {
  const long long p = (long long)a * (long long)b;
  const i32 r = (i32)(p >> 31);
  return ((a & b) == (i32)0x80000000 && a == b) ? 0x7fffffff : r;
}
XB_DEV i32 mul_sh1(i32 a, i32 b) { return (i32)(((long long)a * (long long)b) >> 31); }  // vec_baisc_ops.h:28
// SAT = false: the stage's input bound proves that no product, add, subtract or doubling of the stage can leave 32 bits
// (bounds at fft_batch), so they run as plain wrapping instructions (1 SASS instruction instead of 5) and give the same words
template <bool SAT> XB_DEV i32 fadd(i32 a, i32 b) { return SAT ? add_sat(a, b) : wadd(a, b); }
template <bool SAT> XB_DEV i32 fsub(i32 a, i32 b) { return SAT ? sub_sat(a, b) : wsub(a, b); }
template <bool SAT> XB_DEV i32 fsh1(i32 a) { return SAT ? add_sat(a, a) : (i32)((u32)a << 1); }  // ixheaac_shl32_sat(a, 1)
template <bool SAT> XB_DEV i32 fmul31(i32 a, i32 b) { return SAT ? mul_sat31(a, b) : mul_sh1(a, b); }
XB_DEV i32 amax(i32 m, i32 v) { return max(m, max(v, ~v)); }  // running bound: |v| <= amax + 1

template <bool SAT> XB_DEV void rot_a(i32 &xr, i32 &xi, i32 wh, i32 wl) {
  const i32 t = fadd<SAT>(fmul31<SAT>(xr, wl), fmul31<SAT>(xi, wh));
  xi = fadd<SAT>(wneg(fmul31<SAT>(xr, wh)), fmul31<SAT>(xi, wl));
  xr = t;
}
template <bool SAT> XB_DEV void rot_b(i32 &xr, i32 &xi, i32 wh, i32 wl) {
  const i32 t = fsub<SAT>(fmul31<SAT>(xr, wh), fmul31<SAT>(xi, wl));
  xi = fadd<SAT>(fmul31<SAT>(xr, wl), fmul31<SAT>(xi, wh));
  xr = t;
}
template <bool SAT> XB_DEV void rot_c(i32 &xr, i32 &xi, i32 wh, i32 wl) {
  const i32 t = wneg(fadd<SAT>(fmul31<SAT>(xr, wl), fmul31<SAT>(xi, wh)));
  xi = fadd<SAT>(wneg(fmul31<SAT>(xr, wh)), fmul31<SAT>(xi, wl));
  xr = t;
}

// fft.c:1995-2012 (alt: :2390-2407).  Legs in, the reference's store order out: o0 = x0, o1 = x2, o2 = x1, o3 = (x3i, x3r).
template <bool SAT>
XB_DEV void bfly4(int2 x0, int2 x1, int2 x2, int2 x3, bool alt, int2 &o0, int2 &o1, int2 &o2, int2 &o3) {
  i32 x0r = fadd<SAT>(x0.x, x2.x), x0i = fadd<SAT>(x0.y, x2.y);
  i32 x2r = fsub<SAT>(x0r, fsh1<SAT>(x2.x)), x2i = fsub<SAT>(x0i, fsh1<SAT>(x2.y));
  i32 x1r = fadd<SAT>(x1.x, x3.x);
  i32 x1i = alt ? fsub<SAT>(x1.y, x3.y) : fadd<SAT>(x1.y, x3.y);
  i32 x3r = fsub<SAT>(x1r, fsh1<SAT>(x3.x));
  i32 x3i = alt ? fadd<SAT>(x1i, fsh1<SAT>(x3.y)) : fsub<SAT>(x1i, fsh1<SAT>(x3.y));
  x0r = fadd<SAT>(x0r, x1r);
  x0i = fadd<SAT>(x0i, x1i);
  x1r = fsub<SAT>(x0r, fsh1<SAT>(x1r));
  x1i = fsub<SAT>(x0i, fsh1<SAT>(x1i));
  x2r = fsub<SAT>(x2r, x3i);
  x2i = fadd<SAT>(x2i, x3r);
  x3i = fadd<SAT>(x2r, fsh1<SAT>(x3i));
  x3r = fsub<SAT>(x2i, fsh1<SAT>(x3r));
  o0 = make_int2(x0r, x0i);
  o1 = make_int2(x2r, x2i);
  o2 = make_int2(x1r, x1i);
  o3 = make_int2(x3i, x3r);
}
XB_DEV i32 amax4(i32 m, int2 a, int2 b, int2 c, int2 d) {
  return amax(amax(amax(amax(amax(amax(amax(amax(m, a.x), a.y), b.x), b.y), c.x), c.y), d.x), d.y);
}

XB_DEV unsigned dig_rev16(unsigned v) {  // fft.c:39-46 without the final shift
  v = ((v & 0x33333333u) << 2) | ((v & ~0x33333333u) >> 2);
  v = ((v & 0x0F0F0F0Fu) << 4) | ((v & ~0x0F0F0F0Fu) >> 4);
  v = ((v & 0x00FF00FFu) << 8) | ((v & ~0x00FF00FFu) >> 8);
  return v;
}

XB_DEV i32 ld2(const i32 *buf, int a, int2 &v) {
  v = *reinterpret_cast<const int2 *>(buf + PA(a));
  return 0;
}
XB_DEV void st2(i32 *buf, int a, int2 v) { *reinterpret_cast<int2 *>(buf + PA(a)) = v; }

// ixheaacd_complex_fft_p2_dec, fft_mode = 1, as a batch of `nblk` transforms of `np` points (nblk * np = 512):
// px (interleaved re, im; already divided by 1 << shift) -> y.  Both buffers are padded (PA).
// Every stage exists twice: with the reference's saturating operations, and with wrapping ones for inputs whose bound B
// (|x| <= B, tracked exactly from stage to stage) proves that the exact value of every intermediate fits 32 bits.  With
// rotated legs bounded by 2B the largest intermediates of a butterfly are x0 + x2 + x1 + x3 <= 7B and the doubled terms
// shl(x1r, 1), shl(x3i, 1) <= 8B: B < 2^28 for the middle stages, B < 2^29 for the first one (no rotation: 4B), B < 2^30 for
// the final radix-2 stage (only the rotation can saturate: 2B).  Products of such values with a twiddle fit too, and the one
// saturating product INT_MIN x INT_MIN needs a data word of INT_MIN.
template <bool SAT>
XB_DEV i32 fft_first(const i32 *px, i32 *y, int np, int lane) {
  const bool p512 = np == 512;
  const int lg_q = p512 ? 7 : 4;  // log2(butterflies per block)
  const int dr_shift = p512 ? 6 : 9;
  i32 m = 0;
  // ---- first radix-4 stage with digit reversal (fft.c:1969-2020) ----
#pragma unroll 1
  for (int t = 0; t < 4; t++) {
    const int q = lane + 32 * t;
    const int blk = q >> lg_q, i = (q & ((1 << lg_q) - 1)) << 2;
    int h2 = (int)(dig_rev16((unsigned)i) >> dr_shift);
    if (p512) h2 = (h2 + 1) & ~1;
    const int base = blk * 2 * np;
    int2 a, b, c, d, o0, o1, o2, o3;
    ld2(px, base + h2, a);
    ld2(px, base + h2 + (np >> 1), b);
    ld2(px, base + h2 + np, c);
    ld2(px, base + h2 + np + (np >> 1), d);
    bfly4<SAT>(a, b, c, d, false, o0, o1, o2, o3);
    m = amax4(m, o0, o1, o2, o3);
    st2(y, base + 2 * i, o0);
    st2(y, base + 2 * i + 2, o1);
    st2(y, base + 2 * i + 4, o2);
    st2(y, base + 2 * i + 6, o3);
  }
  return m;
}
// ---- one middle radix-4 stage (fft.c:2025-2422) ----
template <bool SAT>
XB_DEV i32 fft_mid(const i32 *tw, i32 *y, int np, int lane, int del, int lg_del, int nodespacing) {
  const int lg_q = np == 512 ? 7 : 4;
  i32 m = 0;
#pragma unroll 1
  for (int t = 0; t < 4; t++) {
    const int q = lane + 32 * t;
    const int blk = q >> lg_q, rem = q & ((1 << lg_q) - 1);
    const int jj = rem & (del - 1), g = rem >> lg_del;
    const int j = jj * nodespacing;
    const int a0 = blk * 2 * np + 2 * jj + 8 * del * g;
    int2 x0, x1, x2, x3, o0, o1, o2, o3;
    ld2(y, a0, x0);
    ld2(y, a0 + 2 * del, x1);
    ld2(y, a0 + 4 * del, x2);
    ld2(y, a0 + 6 * del, x3);
    bool alt = false;
    if (jj > 0) {
      const i32 w1h = (*(tw + 2 * j)), w1l = (*(tw + 2 * j + 1));
      rot_a<SAT>(x1.x, x1.y, w1h, w1l);
      if (j <= 85) {
        rot_a<SAT>(x2.x, x2.y, (*(tw + 4 * j)), (*(tw + 4 * j + 1)));
        rot_a<SAT>(x3.x, x3.y, (*(tw + 6 * j)), (*(tw + 6 * j + 1)));
      } else if (j <= 128) {
        rot_a<SAT>(x2.x, x2.y, (*(tw + 4 * j)), (*(tw + 4 * j + 1)));
        rot_b<SAT>(x3.x, x3.y, (*(tw + 6 * j - 512)), (*(tw + 6 * j - 511)));
      } else if (j <= 170) {
        rot_b<SAT>(x2.x, x2.y, (*(tw + 4 * j - 512)), (*(tw + 4 * j - 511)));
        rot_b<SAT>(x3.x, x3.y, (*(tw + 6 * j - 512)), (*(tw + 6 * j - 511)));
      } else {
        rot_b<SAT>(x2.x, x2.y, (*(tw + 4 * j - 512)), (*(tw + 4 * j - 511)));
        rot_c<SAT>(x3.x, x3.y, (*(tw + 6 * j - 1024)), (*(tw + 6 * j - 1023)));
        alt = true;
      }
    }
    bfly4<SAT>(x0, x1, x2, x3, alt, o0, o1, o2, o3);
    m = amax4(m, o0, o1, o2, o3);
    st2(y, a0, o0);
    st2(y, a0 + 2 * del, o1);
    st2(y, a0 + 4 * del, o2);
    st2(y, a0 + 6 * del, o3);
  }
  return m;
}
// ---- final radix-2 stage of the 512-point transform (fft.c:2423-2481): del = 256, twiddle step 4 words ----
template <bool SAT>
XB_DEV void fft_last(const i32 *tw, i32 *y, int lane) {
#pragma unroll 1
  for (int t = 0; t < 8; t++) {
    const int q = lane + 32 * t;  // 0..255: first half rot_a (q < 128), second half rot_b
    const int tt = q & 127;
    int2 x0, x1;
    ld2(y, 2 * q, x0);
    ld2(y, 2 * q + 512, x1);
    const i32 wh = (*(tw + 4 * tt)), wl = (*(tw + 4 * tt + 1));
    if (q < 128) rot_a<SAT>(x1.x, x1.y, wh, wl);
    else rot_b<SAT>(x1.x, x1.y, wh, wl);
    st2(y, 2 * q + 512, make_int2(wsub(x0.x / 2, x1.x / 2), wsub(x0.y / 2, x1.y / 2)));
    st2(y, 2 * q, make_int2(wadd(x0.x / 2, x1.x / 2), wadd(x0.y / 2, x1.y / 2)));
  }
}
// in_bound: amax over the words of px (warp-uniform)
XB_DEV void fft_batch(const i32 *tw, const i32 *px, i32 *y, int np, int lane, i32 in_bound) {
  const bool p512 = np == 512;
  i32 m = in_bound < (1 << 29) - 2 ? fft_first<false>(px, y, np, lane) : fft_first<true>(px, y, np, lane);
  m = __reduce_max_sync(0xffffffffu, m);
  __syncwarp();
  const int n_mid = p512 ? 3 : 2;
  int del = 4, lg_del = 2, nodespacing = 64;
#pragma unroll 1
  for (int st = 0; st < n_mid; st++) {
    m = m < (1 << 28) - 2 ? fft_mid<false>(tw, y, np, lane, del, lg_del, nodespacing)
                          : fft_mid<true>(tw, y, np, lane, del, lg_del, nodespacing);
    m = __reduce_max_sync(0xffffffffu, m);
    __syncwarp();
    nodespacing >>= 2;
    del <<= 2;
    lg_del += 2;
  }
  if (p512) {
    if (m < (1 << 30) - 2) fft_last<false>(tw, y, lane);
    else fft_last<true>(tw, y, lane);
    __syncwarp();
  }
}

XB_DEV i32 abs_sat_(i32 a) { return a == (i32)0x80000000 ? 0x7fffffff : (a < 0 ? -a : a); }

// imdct.c:81-91 over the padded buffer (1024 words)
XB_DEV int max_headroom(const i32 *A, int lane) {
  i32 m = 0;
#pragma unroll 4
  for (int i = lane; i < 1024; i += 32) m = max(m, abs_sat_(A[PA(i)]));
  m = __reduce_max_sync(0xffffffffu, m);
  return norm32(m);
}

// ixheaacd_acelp_imdct for a batch: nblk blocks of N = 1024 / nblk coefficients in A (already normalised by `pre_sh`
// on the fly), result back in A.  Returns preshift + 2 (the amount the caller subtracts from its Q).
XB_DEV int imdct_batch(const uint8_t *rom, i32 *A, i32 *B, int nblk, int pre_sh, int lane) {
  const int N = 1024 / nblk, nl = N >> 1;
  const i32 *tw = reinterpret_cast<const i32 *>(rom + kURomFftTw);
  const i32 *cs = reinterpret_cast<const i32 *>(rom + (nl == 512 ? kURomCos512 : kURomCos64));
  const i32 *sn = reinterpret_cast<const i32 *>(rom + (nl == 512 ? kURomSin512 : kURomSin64));
  const int lg = nl == 512 ? 9 : 6;                 // n of fft.c:1431-1441
  const int shift = (lg & 1) ? (lg + 3) / 2 : (lg + 4) / 2;
  const int div = 1 << shift;
  // pre-twiddle (imdct.c:111-127) in place, pairs (i, nl-1-i): complex i at words 2i, 2i+1 of its block
  i32 bound = 0;
#pragma unroll 2
  for (int q = lane; q < 256; q += 32) {
    const int ppb = nl >> 1;  // pairs per block
    const int blk = q / ppb, i = q % ppb, i2 = nl - 1 - i;
    const int base = blk * N;
    int2 lo, hi;
    ld2(A, base + 2 * i, lo);       // x[2i], x[2i+1]
    ld2(A, base + 2 * i2, hi);      // x[2 i2], x[2 i2 + 1] = x[2nl-1-2i]
    lo.x = lsl(lo.x, pre_sh); lo.y = lsl(lo.y, pre_sh); hi.x = lsl(hi.x, pre_sh); hi.y = lsl(hi.y, pre_sh);
    const i32 c1 = (*(cs + i)), s1 = (*(sn + i)), c2 = (*(cs + i2)), s2 = (*(sn + i2));
    const i32 r1 = wsub(mul32(neg_sat(lo.x), c1), mul32(hi.y, s1));
    const i32 m1 = wsub(mul32(hi.y, c1), mul32(lo.x, s1));
    const i32 r2 = wsub(mul32(neg_sat(hi.x), c2), mul32(lo.y, s2));
    const i32 m2 = wsub(mul32(lo.y, c2), mul32(hi.x, s2));
    const int2 v1 = make_int2(r1 / div, m1 / div), v2 = make_int2(r2 / div, m2 / div);  // fft.c:1443-1446: C division
    bound = amax(amax(amax(amax(bound, v1.x), v1.y), v2.x), v2.y);
    st2(A, base + 2 * i, v1);
    st2(A, base + 2 * i2, v2);
  }
  bound = __reduce_max_sync(0xffffffffu, bound);
  __syncwarp();
  fft_batch(tw, A, B, nl, lane, bound);
  // post-twiddle (imdct.c:129-147): B (r, im interleaved) -> A
#pragma unroll 2
  for (int q = lane; q < 512; q += 32) {
    const int blk = q / nl, i = q % nl;
    int2 v;
    ld2(B, blk * N + 2 * i, v);
    const i32 c = (*(cs + i)), s = (*(sn + i));
    A[PA(blk * N + 2 * i)] = wneg(wsub(mul32(v.x, c), mul32(v.y, s)));
    A[PA(blk * N + 2 * nl - 1 - 2 * i)] = wneg(wadd(mul32(v.y, c), mul32(v.x, s)));
  }
  __syncwarp();
  int preshift = (nl == 512) ? 10 : 7;  // imdct.c:190-195 for N = 1024 / 128
  preshift = (shift + (nl == 512 ? 1 : 0)) - preshift;
  return preshift + 2;
}

__global__ void __launch_bounds__(kUfWarps * 32) usac_fd_kernel(UsacFdArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // the tables (15.9 KB: FFT twiddles, pre / post twiddles, windows) first, then the per-warp buffers
  const uint8_t *rom = smem_raw;
  {
    const i32 *src = reinterpret_cast<const i32 *>(p.rom);
    i32 *dst = reinterpret_cast<i32 *>(smem_raw);
    for (int i = threadIdx.x; i < kURomBytes / 4; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  i32 *smem = reinterpret_cast<i32 *>(smem_raw + kURomPad);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  i32 *A = smem + warp * 2 * kUfBuf, *B = A + kUfBuf;
  const long long warps_total = (long long)gridDim.x * kUfWarps;
  const int so = 14;
  for (long long u = (long long)blockIdx.x * kUfWarps + warp; u < p.n_units; u += warps_total) {
    __syncwarp();
    const int win_seq = p.ics[2 * u], win_shape = p.ics[2 * u + 1], shape_prev = p.wstate[u];
    if (u + warps_total < p.n_units) {  // pull this warp's next unit (coefficients + overlap, 8 KB) towards L2
      const char *q0 = reinterpret_cast<const char *>(p.coef + (u + warps_total) * 1024);
      const char *q1 = reinterpret_cast<const char *>(p.overlap + (u + warps_total) * 1024);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(q0 + lane * 128));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(q1 + lane * 128));
    }
    {
      const int4 *src = reinterpret_cast<const int4 *>(p.coef + u * 1024);
      int4 v[8];  // the unit's 4 KB of coefficients: 8 x 16-byte requests per lane in flight
#pragma unroll
      for (int q = 0; q < 8; q++) v[q] = __ldg(src + lane + 32 * q);
#pragma unroll
      for (int q = 0; q < 8; q++) *reinterpret_cast<int4 *>(A + PA(4 * (lane + 32 * q))) = v[q];
    }
    __syncwarp();
    i32 *ov = p.overlap + u * 1024, *out = p.out + u * 1024;
    int max_shift = max_headroom(A, lane);
    int shiftp = (int)(int8_t)(max_shift + 6);
    const bool is_short = win_seq == 2;
    shiftp = (int)(int8_t)(shiftp - imdct_batch(rom, A, B, is_short ? 8 : 1, max_shift, lane));
    max_shift = max_headroom(A, lane);
    // imdct.c:93-99 with max_shift - 1: a count of -1 clears the block in the reference build (PSLLD, see oracle)
    const int nsh = max_shift - 1;
    shiftp = (int)(int8_t)(shiftp + nsh);
    if (shiftp - so > 31) shiftp = 31 + so;
    auto IN = [&](int i) -> i32 { return nsh < 0 ? 0 : lsl(A[PA(i)], nsh); };

    if (!is_short) {
      int output_q;
      if (win_seq == 0 || win_seq == 1) {  // ixheaacd_windowing_long1 (basic_ops.c:77-123)
        const i32 *win = reinterpret_cast<const i32 *>(rom + (shape_prev ? kURomKbd1024 : kURomSine1024));
        const bool gt = shiftp > so;
        const int sh = gt ? shiftp - so : so - shiftp;
#pragma unroll 2
        for (int i = lane; i < 512; i += 32) {
          const i32 wf = (*(win + i)), wr = (*(win + 1023 - i)), a = IN(512 + i), o1 = ov[i], o2 = ov[1023 - i];
          i32 d1, d2;
          if (gt) {
            d1 = add_sat(mul_sh1(a, wf) >> sh, mul_sh1(o1, wr));
            d2 = add_sat(mul_sh1(neg_sat(a), wr) >> sh, mul_sh1(o2, wf));
          } else {
            d1 = add_sat(mul_sh1(a, wf), mul_sh1(o1, wr) >> sh);
            d2 = add_sat(mul_sh1(neg_sat(a), wr), mul_sh1(o2, wf) >> sh);
          }
          B[PA(i)] = d1;
          B[PA(1023 - i)] = d2;
        }
        output_q = gt ? so : shiftp;
      } else {  // ixheaacd_windowing_long3 (basic_ops.c:298-372), n_flat = 448, n_trans = 128
        const i32 *wsh = reinterpret_cast<const i32 *>(rom + (shape_prev ? kURomKbd128 : kURomSine128));
        const bool gt = shiftp > so;
        const int sh = gt ? shiftp - so : so - shiftp;
#pragma unroll 2
        for (int i = lane; i < 1024; i += 32) {
          i32 d;
          if (i < 448) d = gt ? ov[i] : (ov[i] >> sh);
          else if (i < 576) {
            const i32 a = i < 512 ? IN(512 + i) : neg_sat(IN(512 + 1023 - i));
            const i32 wf = (*(wsh + i - 448)), wr = (*(wsh + 127 - (i - 448)));
            d = gt ? add_sat(mul_sh1(a, wf) >> sh, mul_sh1(ov[i], wr)) : add_sat(mul_sh1(a, wf), mul_sh1(ov[i], wr) >> sh);
          } else {
            const i32 a = neg_sat(IN(512 + 1023 - i));
            d = gt ? (a >> sh) : a;
          }
          B[PA(i)] = d;
        }
        output_q = gt ? so : shiftp;
      }
      __syncwarp();
      {  // overlap for the next frame (imdct.c:553-568) and ixheaacd_scale_down_adj(.., output_q, 15) (basic_ops.c:641-657)
        const int sh = shiftp > so ? shiftp - so : so - shiftp;
#pragma unroll 2
        for (int i = lane; i < 512; i += 32) {
          const i32 v = neg_sat(IN(i)) >> sh;
          ov[512 + i] = v;
          ov[511 - i] = v;
        }
#pragma unroll 2
        for (int i = lane; i < 1024; i += 32) {
          const i32 d = B[PA(i)];
          out[i] = add_sat(output_q > 15 ? (d >> (output_q - 15)) : shl32_sat(d, 15 - output_q), 11);
        }
      }
    } else {
      // EIGHT_SHORT (imdct.c:336-475): the 2048-word work buffer is out (low half) | ov (high half) in HBM / L2
      const i32 *wsh = reinterpret_cast<const i32 *>(rom + (win_shape ? kURomKbd128 : kURomSine128));
      const i32 *wpv = reinterpret_cast<const i32 *>(rom + (shape_prev ? kURomKbd128 : kURomSine128));
      auto BUF = [&](int i) -> i32 * { return i < 1024 ? out + i : ov + (i - 1024); };
      for (int i = lane; i < 1024; i += 32) {
        out[i] = ov[i];
      }
      __syncwarp();
      for (int i = lane; i < 1024; i += 32) ov[i] = 0;
      __syncwarp();
      const bool so_gt = so > shiftp;
      {  // ixheaacd_windowing_short2 (basic_ops.c:429-478): src1 = in + 64, fp = buf + 448
        const int sh = so_gt ? so - shiftp : shiftp - so;
        for (int i = lane; i < 64; i += 32) {
          const i32 wf = (*(wpv + i)), wr = (*(wpv + 127 - i)), a = IN(64 + i);
          i32 *f1 = BUF(448 + i), *f2 = BUF(448 + 127 - i);
          if (so_gt) {
            *f1 = add_sat(mul_sh1(a, wf), mul_sh1(*f1, wr) >> sh);
            *f2 = add_sat(mul_sh1(neg_sat(a), wr), mul_sh1(*f2, wf) >> sh);
          } else {
            *f1 = add_sat(mul_sh1(a, wf) >> sh, mul_sh1(*f1, wr));
            *f2 = add_sat(mul_sh1(neg_sat(a), wr) >> sh, mul_sh1(*f2, wf));
          }
        }
        for (int i = 128 + lane; i < 448 + 128; i += 32) *BUF(448 + i) = 0;
      }
      __syncwarp();
      const int oq = so_gt ? shiftp : so;
      {  // ixheaacd_windowing_short3 (basic_ops.c:480-521): src1 = in, fp = buf + 576
        const int sh = so_gt ? so - shiftp : shiftp - so;
        for (int i = lane; i < 64; i += 32) {
          const i32 wr = (*(wsh + 127 - i)), wf = (*(wsh + i)), a = neg_sat(IN(63 - i));
          i32 *f1 = BUF(576 + i), *f2 = BUF(576 + 127 - i);
          if (so_gt) {
            *f1 = add_sat(mul_sh1(a, wr), *f1 >> sh);
            *f2 = add_sat(mul_sh1(a, wf), *f2 >> sh);
          } else {
            *f1 = add_sat(mul_sh1(a, wr) >> sh, *f1);
            *f2 = add_sat(mul_sh1(a, wf) >> sh, *f2);
          }
        }
      }
      __syncwarp();
      // ixheaacd_windowing_short4 x 7 (basic_ops.c:523-621): block k = 1..7, src1 = in + 128 k, fp = buf + 448 + 128 k
      const bool big = so > oq;
      const int sh4 = big ? shiftp - oq : shiftp - so;
#pragma unroll 1
      for (int k = 1; k < 8; k++) {
        const int fp0 = 448 + 128 * k, s0 = 128 * k;
        const bool flag = k < 7;
        for (int i = lane; i < 64; i += 32) {
          const i32 wf = (*(wsh + i)), wr = (*(wsh + 127 - i)), a = IN(s0 + 64 + i);
          i32 *f1 = BUF(fp0 + i), *f2 = BUF(fp0 + 127 - i);
          if (big) {
            *f1 = add_sat(mul_sh1(a, wf) >> sh4, *f1);
            *f2 = add_sat(mul_sh1(neg_sat(a), wr) >> sh4, *f2);
          } else {
            *f1 = add_sat(mul_sh1(a, wf) >> sh4, *f1 >> (oq - so));
            *f2 = add_sat(mul_sh1(neg_sat(a), wr) >> sh4, *f2);
          }
        }
        __syncwarp();
        for (int i = 64 + lane; i < 128; i += 32) {
          const int t = i - 64;
          const i32 a = neg_sat(IN(s0 + 127 - i));
          i32 *pa = BUF(fp0 + i + 64), *pb = BUF(fp0 + 384 - 64 - i - 1);
          const i32 va = flag ? mul_sh1(a, (*(wsh + 127 - t))) : a, vb = flag ? mul_sh1(a, (*(wsh + t))) : a;
          if (big) {
            *pa = add_sat(va >> sh4, *pa >> (so - oq));
            *pb = add_sat(vb >> sh4, *pb >> (so - oq));
          } else {
            *pa = add_sat(va >> sh4, *pa);
            *pb = add_sat(vb >> sh4, *pb);
          }
        }
        __syncwarp();
      }
      // imdct.c:441-451: clear the tail, rescale: buf[0..447] so -> oq; ov <- buf[1024..] oq -> so; out <- buf[0..1023] oq -> 15
      for (int i = 2048 - 448 + lane; i < 2048; i += 32) *BUF(i) = 0;
      for (int i = lane; i < 448; i += 32) out[i] = so > oq ? (out[i] >> (so - oq)) : shl32_sat(out[i], oq - so);
      __syncwarp();
      for (int i = lane; i < 1024; i += 32) {
        const i32 h = ov[i], l = out[i];
        ov[i] = oq > so ? (h >> (oq - so)) : shl32_sat(h, so - oq);
        out[i] = oq > 15 ? (l >> (oq - 15)) : shl32_sat(l, 15 - oq);
      }
    }
    if (lane == 0) p.wstate[u] = (uint8_t)win_shape;  // window_shape_prev = window_shape (ext_ch_ele.c:968)
  }
}

size_t usac_fd_smem_bytes() { return (size_t)kURomPad + (size_t)kUfWarps * 2 * kUfBuf * 4; }

// nothing in the kernel depends on table values beyond their layout; kept as the install-time hook
int usac_fd_check_tables(const uint8_t *urom) { return urom ? 0 : -1; }

cudaError_t launch_usac_fd(const UsacFdArgs &args, int num_sms, cudaStream_t stream) {
  static xb::PerDeviceOnce configured;
  const size_t smem = usac_fd_smem_bytes();
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(usac_fd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.done();
  }
  long long need = (args.n_units + kUfWarps - 1) / kUfWarps;
  long long grid = (long long)num_sms * 2;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  usac_fd_kernel<<<(unsigned)grid, kUfWarps * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace xb
