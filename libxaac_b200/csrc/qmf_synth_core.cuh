// qmf_synth_core.cuh — the arithmetic of the fixed-point HQ 64-band QMF synthesis, written per LANE so that the
// same source runs inside qmf_synth_hq_kernel (device) and inside the CPU lane simulator the tests use to check the
// index maps without a GPU (tests/sim/synth_sim.cu; test infrastructure, never linked into the product library).
//
// Work split of one unit (32 slots x 64 complex bands -> 2048 PCM16 samples), one warp per unit:
//   phase A  lane = SLOT.  The lane owns one row of the matrix (128 words, staged in shared memory by a 512-byte
//            bulk copy), block-shifts it, and runs the whole modulation of that slot in registers:
//            pre-twiddle -> two radix-4 stages -> radix-2 + digit reversal -> post-twiddle -> fold to 128 filter-state
//            samples.  No shuffles, no shared-memory exchanges, all twiddles are constant-bank operands.  The 128
//            samples overwrite the lane's own row (sign-extended, one per word, pair-interleaved: word
//            4k'+2h+e = sample 64h+2k'+e, so that phase B reads both 64-sample halves of a pair with one LDS.128).
//   phase B  lane = OUTPUT PAIR k' (samples 2k', 2k'+1 of every slot).  The ring of the reference
//            (filter_states[1280], write offset moving backwards by 128 per slot) is restated in linear time: row r
//            holds the block folded at slot r, rows -9..-1 hold the nine blocks of history, and the 10-tap window of
//            slot s is sum_a row[s-a][half(a)] * w_a with per-unit register coefficients — see window_group().
//
// Reference (paths relative to /root/reference):
//   ixheaacd_cplx_synt_qmffilt            decoder/ixheaacd_qmf_dec.c:811-1129
//   ixheaacd_adjust_scale_dec             decoder/ixheaacd_env_calc.c:1099-1157
//   ixheaacd_cos_sin_mod                  decoder/generic/ixheaacd_qmf_dec_generic.c:259-466
//   ixheaacd_radix4bfly                   generic:1736-1829
//   ixheaacd_postradixcompute2            generic:1934-2015
//   ixheaacd_shiftrountine_with_rnd       generic:1638-1670
//   ixheaacd_sbr_qmfsyn64_winadd          generic:1508-1542
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace xb {
namespace syn {

typedef int32_t i32;
typedef uint32_t u32;

#define XB_HD __host__ __device__ __forceinline__

constexpr int kRowW = 132;   // words per row in shared memory (128 used; 132 = 4 mod 32 keeps LDS.128 by row conflict-free)
constexpr int kHist = 9;     // rows of history in front of the 32 rows of the frame
constexpr int kRows = kHist + 32;

// twiddles, every entry (first << 16, second << 16) of the reference's WORD16 pairs
struct SynTw {
  int2 pre[32];  // sbr_sin_cos_twiddle_l64: (wim, wre)
  int2 alt[16];  // sbr_alt_sin_twiddle_l64: (wim, wre)
  int2 w1[24];   // w_32, radix-4 stage 1: position i -> (si, co) x 3
  int2 w2[6];    // w_32 + 48, radix-4 stage 2
};

XB_HD i32 mh(i32 a, i32 b) {
#ifdef __CUDA_ARCH__
  return __mulhi(a, b);
#else
  return (i32)(((int64_t)a * (int64_t)b) >> 32);
#endif
}
XB_HD i32 sat_add(i32 a, i32 b) {
#ifdef __CUDA_ARCH__
  i32 r;
  asm("add.sat.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
#else
  int64_t s = (int64_t)a + b;
  return (i32)(s > 2147483647LL ? 2147483647LL : (s < -2147483648LL ? -2147483648LL : s));
#endif
}
XB_HD i32 sat_sub(i32 a, i32 b) {
#ifdef __CUDA_ARCH__
  i32 r;
  asm("sub.sat.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
#else
  int64_t s = (int64_t)a - b;
  return (i32)(s > 2147483647LL ? 2147483647LL : (s < -2147483648LL ? -2147483648LL : s));
#endif
}
XB_HD i32 imin(i32 a, i32 b) { return a < b ? a : b; }
XB_HD i32 imax(i32 a, i32 b) { return a > b ? a : b; }

// SAT = false: wrapping adds, bit-identical to the saturating ones as long as the unit's block-shifted inputs stay
// below 2^fast_bits (bound derived from the tables, qmf_synth_build_tables); SAT = true: the reference's adds.
// The wrapping forms carry a run-time zero `z` as third addend where the sum feeds on from IMAD.HI results: that keeps
// the add on the ALU pipe (IADD3) instead of being folded into the multiplier's 64-bit addend.
template <bool SAT> XB_HD i32 A_(i32 a, i32 b) { return SAT ? sat_add(a, b) : (i32)((u32)a + (u32)b); }
template <bool SAT> XB_HD i32 S_(i32 a, i32 b) { return SAT ? sat_sub(a, b) : (i32)((u32)a - (u32)b); }
template <bool SAT> XB_HD i32 N_(i32 a) { return SAT ? sat_sub(0, a) : (i32)(0u - (u32)a); }
template <bool SAT> XB_HD i32 RA_(i32 a, i32 b, i32 z) { return SAT ? sat_add(a, b) : (i32)((u32)a + (u32)b + (u32)z); }
template <bool SAT> XB_HD i32 RS_(i32 a, i32 b, i32 z) { return SAT ? sat_sub(a, b) : (i32)((u32)a - (u32)b + (u32)z); }
XB_HD i32 add3(i32 a, i32 b, i32 z) { return (i32)((u32)a + (u32)b + (u32)z); }
XB_HD i32 sub3(i32 a, i32 b, i32 z) { return (i32)((u32)a - (u32)b + (u32)z); }
XB_HD i32 shl1(i32 a) { return (i32)((u32)a << 1); }

// one radix-4 butterfly, generic:1766-1822; t1..t3 = (si << 16, co << 16)
template <bool SAT>
XB_HD void radix4(int2 &e0, int2 &e1, int2 &e2, int2 &e3, const int2 t1, const int2 t2, const int2 t3, const i32 z) {
  i32 xh0 = A_<SAT>(e0.x, e2.x), xl0 = S_<SAT>(e0.x, e2.x);
  i32 xh20 = A_<SAT>(e1.x, e3.x), xl20 = S_<SAT>(e1.x, e3.x);
  i32 xh1 = A_<SAT>(e0.y, e2.y), xl1 = S_<SAT>(e0.y, e2.y);
  i32 xh21 = A_<SAT>(e1.y, e3.y), xl21 = S_<SAT>(e1.y, e3.y);
  i32 xt0 = S_<SAT>(xh0, xh20), yt0 = S_<SAT>(xh1, xh21);
  i32 xt1 = A_<SAT>(xl0, xl21), xt2 = S_<SAT>(xl0, xl21);
  i32 yt2 = A_<SAT>(xl1, xl20), yt1 = S_<SAT>(xl1, xl20);
  e0.x = A_<SAT>(xh0, xh20);
  e0.y = A_<SAT>(xh1, xh21);
  e3.x = shl1(add3(mh(yt2, t3.x), mh(xt2, t3.y), z));
  e3.y = shl1(sub3(mh(yt2, t3.y), mh(xt2, t3.x), z));
  e2.x = shl1(add3(mh(yt0, t2.x), mh(xt0, t2.y), z));
  e2.y = shl1(sub3(mh(yt0, t2.y), mh(xt0, t2.x), z));
  e1.x = shl1(add3(mh(yt1, t1.x), mh(xt1, t1.y), z));
  e1.y = shl1(sub3(mh(yt1, t1.y), mh(xt1, t1.x), z));
}

// Pre-twiddle (generic:290-367) of one half of a slot into the FFT's natural positions.  x[64]: the block-shifted real
// parts (HALF 0) or imaginary parts (HALF 1) of the 64 bands.  Step n pairs band n with band 63-n; odd steps are the even
// formula with the two operands swapped and fill from the back.
template <bool SAT, int HALF>
XB_HD void pre_twiddle(const i32 (&x)[64], int2 (&E)[32], const SynTw &tw, const i32 z) {
#pragma unroll
  for (int n = 0; n < 32; n++) {
    const int pos = (n & 1) ? 31 - (n >> 1) : (n >> 1);
    const i32 a = (n & 1) ? x[63 - n] : x[n];
    const i32 b = (n & 1) ? x[n] : x[63 - n];
    const i32 wim = tw.pre[n].x, wre = tw.pre[n].y;
    if (HALF == 0) {
      E[pos].x = RA_<SAT>(mh(a, wre), mh(b, wim), z);
      E[pos].y = RS_<SAT>(mh(b, wre), mh(a, wim), z);
    } else {  // (c, d) = (a, b)
      E[pos].x = RS_<SAT>(mh(b, wim), mh(a, wre), z);
      E[pos].y = RA_<SAT>(mh(a, wim), mh(b, wre), z);
    }
  }
}
// the two radix-4 stages of the 32-point FFT (generic:369-386), in registers
template <bool SAT>
XB_HD void fft_stages(int2 (&E)[32], const SynTw &tw, const i32 z) {
#pragma unroll
  for (int i = 0; i < 8; i++)
    radix4<SAT>(E[i], E[i + 8], E[i + 16], E[i + 24], tw.w1[3 * i], tw.w1[3 * i + 1], tw.w1[3 * i + 2], z);
#pragma unroll
  for (int g = 0; g < 4; g++)
#pragma unroll
    for (int i = 0; i < 2; i++)
      radix4<SAT>(E[8 * g + i], E[8 * g + i + 2], E[8 * g + i + 4], E[8 * g + i + 6], tw.w2[3 * i], tw.w2[3 * i + 1],
                  tw.w2[3 * i + 2], z);
}

// round16(shl32_sat(x, out_shift)) (generic:1638-1670) == ((clamp(x) << s) + 0x8000) >> 16 with these bounds
struct FoldK {
  i32 lo, hi, mul;
};
XB_HD FoldK fold_consts(int out_shift) {
  FoldK k;
  k.lo = (i32)0x80000000 >> out_shift;
  k.hi = (i32)(0x7fff7fffu >> out_shift);
  k.mul = (i32)(1u << out_shift);
  return k;
}
XB_HD i32 fold_rnd(i32 x, const FoldK &k) {
  x = imax(k.lo, imin(k.hi, x));
  return (i32)((u32)x * (u32)k.mul + 0x8000u) >> 16;
}

// Radix-2 + digit reversal (generic:1934): F[p] = X[src(p)] + X[src(p) + 1] for p < 16, X[src(p-16)] - X[src(p-16) + 1]
// above, src(p) = 8 (p & 3) + 2 (p >> 2) — the standard dig_rev_table2_32, checked when the ROM is installed.
XB_HD int post_src(int p) { return 8 * (p & 3) + 2 * ((p >> 2) & 3); }
// Where the 4-word chunk of output pair c (both 64-sample halves: st[2c], st[2c+1], st[64+2c], st[65+2c]) lives inside a
// row: the fold writes every chunk over FFT words it has just consumed (see post_fold), which permutes the chunks; the
// window, the history and the state save go through the same map.
XB_HD int chunk_pos(int c) { return 16 * (c >> 4) + 4 * (c & 3) + ((c >> 2) & 3); }

// Post-twiddle (generic:388-465) and fold (generic:1638) of output pairs q and 31-q of one slot from the radix-2 outputs
// F[q] ("front", both halves: Ff0 / Ff1) and F[31-q] ("back"), alt_f = alt[q-1], alt_b = alt[q]; q = 0 is the special
// first pair.  lo = chunk q, hi = chunk 31-q.
template <bool SAT>
XB_HD void fold_pair(const int2 Ff0, const int2 Ff1, const int2 Fb0, const int2 Fb1, const int2 alt_f, const int2 alt_b,
                     const bool first, const FoldK &fk, const i32 z, int4 &lo, int4 &hi) {
  // ---- j = 2q (even): s1[j], s2[j], s1[63-j], s2[63-j]
  i32 s1j, s2j, s1m, s2m;
  if (first) {
    s1j = Ff0.x >> 1;
    s1m = N_<SAT>(Ff0.y >> 1);
    s2m = N_<SAT>(Ff1.x >> 1);
    s2j = Ff1.y >> 1;
  } else {
    const i32 wim = alt_f.x, wre = alt_f.y;
    i32 fim = Ff0.x, fre = Ff0.y;
    s1j = RA_<SAT>(mh(fre, wim), mh(fim, wre), z);
    s1m = RS_<SAT>(mh(fim, wim), mh(fre, wre), z);
    fim = Ff1.x;
    fre = Ff1.y;
    s2m = N_<SAT>(RA_<SAT>(mh(fre, wim), mh(fim, wre), z));
    s2j = RS_<SAT>(mh(fre, wre), mh(fim, wim), z);
  }
  // ---- j = 2q+1 (odd): s1[2q+1], s2[2q+1], s1[62-2q], s2[62-2q]
  i32 t1j, t2j, t1m, t2m;
  {
    const i32 wim = alt_b.x, wre = alt_b.y;
    i32 re = Fb0.y, im = Fb0.x;
    t1m = RA_<SAT>(mh(re, wre), mh(im, wim), z);
    t1j = RS_<SAT>(mh(im, wre), mh(re, wim), z);
    re = Fb1.y;
    im = Fb1.x;
    t2j = N_<SAT>(RA_<SAT>(mh(re, wre), mh(im, wim), z));
    t2m = RS_<SAT>(mh(re, wim), mh(im, wre), z);
  }
  // fold: j: r1 = s1[j], i1 = s2[j], r2 = s1[63-j], i2 = s2[63-j]
  //   st[j] = R(i1 - r1)  st[64+j] = R(i2 + r2)  st[63-j] = R(i2 - r2)  st[127-j] = R(i1 + r1)
  lo.x = fold_rnd(S_<SAT>(s2j, s1j), fk);  // st[2q]
  lo.y = fold_rnd(S_<SAT>(t2j, t1j), fk);  // st[2q+1]
  lo.z = fold_rnd(A_<SAT>(s2m, s1m), fk);  // st[64+2q]
  lo.w = fold_rnd(A_<SAT>(t2m, t1m), fk);  // st[65+2q]
  hi.x = fold_rnd(S_<SAT>(t2m, t1m), fk);  // st[62-2q]
  hi.y = fold_rnd(S_<SAT>(s2m, s1m), fk);  // st[63-2q]
  hi.z = fold_rnd(A_<SAT>(t2j, t1j), fk);  // st[126-2q]
  hi.w = fold_rnd(A_<SAT>(s2j, s1j), fk);  // st[127-2q]
}

template <bool SAT> XB_HD int2 r2_sum(const int4 v) { return make_int2(A_<SAT>(v.x, v.z), A_<SAT>(v.y, v.w)); }
template <bool SAT> XB_HD int2 r2_dif(const int4 v) { return make_int2(S_<SAT>(v.x, v.z), S_<SAT>(v.y, v.w)); }

// Radix-2, post-twiddle and fold of one slot, in place.  On entry the row holds the two half FFTs before the radix-2:
// chunk c (4 words) = (X0[2c], X0[2c+1]) for c < 16, (X1[2(c-16)], X1[2(c-16)+1]) above.  Step j handles output pairs
// j, 31-j, 15-j and 16+j: they need exactly the source chunks A = src(j)/2 and B = src(15-j)/2 of both halves (sum and
// difference of one chunk are F[p] and F[p+16]) and their four result chunks go back to those four places — chunk c ends
// up at chunk_pos(c).  A rolled loop: the instruction footprint of the whole kernel has to fit the instruction cache.
template <bool SAT>
XB_HD void post_fold(int4 *r4, const SynTw &tw, const FoldK &fk, const i32 z) {
#pragma unroll 1
  for (int j = 0; j < 8; j++) {
    const int k = 15 - j;
    const int A = 4 * (j & 3) + (j >> 2), B = 4 * (k & 3) + (k >> 2);
    const int4 a0 = r4[A], b0 = r4[B], a1 = r4[16 + A], b1 = r4[16 + B];
    int4 lo, hi, lo2, hi2;
    // pairs j / 31-j: front F[j] = sum(A), back F[31-j] = dif(B)
    fold_pair<SAT>(r2_sum<SAT>(a0), r2_sum<SAT>(a1), r2_dif<SAT>(b0), r2_dif<SAT>(b1), tw.alt[j > 0 ? j - 1 : 0], tw.alt[j],
                   j == 0, fk, z, lo, hi);
    // pairs 15-j / 16+j: front F[15-j] = sum(B), back F[16+j] = dif(A)
    fold_pair<SAT>(r2_sum<SAT>(b0), r2_sum<SAT>(b1), r2_dif<SAT>(a0), r2_dif<SAT>(a1), tw.alt[k - 1], tw.alt[k], false, fk, z,
                   lo2, hi2);
    r4[A] = lo;        // chunk j      -> chunk_pos(j)      = A
    r4[16 + B] = hi;   // chunk 31-j   -> chunk_pos(31-j)   = 16 + B
    r4[B] = lo2;       // chunk 15-j   -> chunk_pos(15-j)   = B
    r4[16 + A] = hi2;  // chunk 16+j   -> chunk_pos(16+j)   = 16 + A
  }
}

// ---- phase B helpers -------------------------------------------------------------------------------------------
// Ring restated in linear time.  The reference writes the block of slot s at ring block Bw_s = (Bw_0 - s) mod 10
// (Bw_0 = drc_offset >> 7) and windows ring block B with coefficient block fp_s + B (fp_s = (filter_pos >> 6) + s
// mod 10) from half (s & 1) ^ (B & 1).  With a = age of a block (0 = the slot's own): B = (Bw_s + a) mod 10, so the
// half is (Bw_0 + a) & 1 and — qmf_c being 640-periodic, checked at ROM install — the coefficient block is
// (fp_0 + Bw_0 + a) mod 10: both depend on the age only.  P = Bw_0 & 1 is a template parameter of the window.
struct WinCoef {
  i32 w[10][2];  // [age][e], sign-extended qmf_c values of this lane's output pair
};
// c32: qmf_c as 32-bit words (two consecutive WORD16 coefficients per word), 640 words
XB_HD void window_coefs(WinCoef &wc, const i32 *c32, int lane, int Bw0, int fp0) {
#pragma unroll
  for (int a = 0; a < 10; a++) {
    int blk = fp0 + Bw0 + a;  // < 30
    blk -= (blk >= 20) ? 20 : (blk >= 10 ? 10 : 0);
    const i32 v = c32[32 * blk + lane];
    wc.w[a][0] = (i32)(int16_t)v;
    wc.w[a][1] = v >> 16;
  }
}

// window of G consecutive slots s0..s0+G-1 for output pair `lane`: rows4 = row -kHist of the warp's buffer viewed as
// int4 (kRowW / 4 per row).  o[t] = PCM16 samples (2 lane, 2 lane + 1) of slot s0 + t, not yet packed.
template <int P, int G>
XB_HD void window_group(const int4 *rows4, int cp, int s0, const WinCoef &wc, i32 (&o)[G][2]) {
  i32 acc[G][2];
#pragma unroll
  for (int t = 0; t < G; t++) acc[t][0] = acc[t][1] = 0x4000;
#pragma unroll
  for (int r = -kHist; r < G; r++) {
    const int4 x = rows4[(s0 + r + kHist) * (kRowW / 4) + cp];
#pragma unroll
    for (int t = 0; t < G; t++) {
      const int a = t - r;
      if (a < 0 || a > 9) continue;
      const bool h = ((P + a) & 1) != 0;
      acc[t][0] += (h ? x.z : x.x) * wc.w[a][0];
      acc[t][1] += (h ? x.w : x.y) * wc.w[a][1];
    }
  }
  // shl32_sat(acc, 1) >> 16  ==  clamp(acc, -2^30, 2^30 - 1) >> 15   (the accumulation itself cannot saturate:
  // sum |c| over the 10 taps <= 57308, checked at ROM install)
#pragma unroll
  for (int t = 0; t < G; t++) {
    o[t][0] = imax(-0x40000000, imin(0x3fffffff, acc[t][0])) >> 15;
    o[t][1] = imax(-0x40000000, imin(0x3fffffff, acc[t][1])) >> 15;
  }
}

// history: ring block B of filter_states (WORD16[1280], words 32 (2B + h) + lane hold samples 128 B + 64 h + 2 lane,
// +1) -> row -a, a = (B - Bw0) mod 10, a != 0.  st32 = the unit's filter_states viewed as 32-bit words.
XB_HD void history_store(int4 *rows4, int lane, int B, int Bw0, i32 w_h0, i32 w_h1) {
  int a = B - Bw0;
  if (a < 0) a += 10;
  if (a == 0) return;  // the block slot 0 overwrites
  rows4[(kHist - a) * (kRowW / 4) + chunk_pos(lane)] = make_int4((i32)(int16_t)w_h0, w_h0 >> 16, (i32)(int16_t)w_h1, w_h1 >> 16);
}
// state save: row r (22..31) -> ring block (Bw0 - r) mod 10
XB_HD void state_words(const int4 *rows4, int lane, int r, i32 &w_h0, i32 &w_h1) {
  const int4 x = rows4[(kHist + r) * (kRowW / 4) + chunk_pos(lane)];
  w_h0 = (i32)(((u32)x.x & 0xffffu) | ((u32)x.y << 16));
  w_h1 = (i32)(((u32)x.z & 0xffffu) | ((u32)x.w << 16));
}

// block shift of one band (env_calc.c:1099): value * mul >> shr with (mul, shr) from shift_entry()
XB_HD int2 shift_entry(int sh) {
  sh = imax(-31, imin(31, sh));
  return make_int2(sh > 0 ? (i32)(1u << sh) : 1, sh < 0 ? -sh : 0);
}
XB_HD i32 shift_val(i32 x, int2 ms) { return (i32)((u32)x * (u32)ms.x) >> ms.y; }

// block shift of the lane's row in place (shv: the 64 (mul, shr) entries of the lane's variant) and the bounds of the result
XB_HD void shift_row(i32 *row, const int2 *shv, i32 &mx, i32 &mn) {
  int4 *r4 = reinterpret_cast<int4 *>(row);
#pragma unroll 4
  for (int q = 0; q < 32; q++) {
    int4 v = r4[q];
    const int4 *e = reinterpret_cast<const int4 *>(shv + 4 * (q & 15));
    const int4 s01 = e[0], s23 = e[1];
    v.x = shift_val(v.x, make_int2(s01.x, s01.y));
    v.y = shift_val(v.y, make_int2(s01.z, s01.w));
    v.z = shift_val(v.z, make_int2(s23.x, s23.y));
    v.w = shift_val(v.w, make_int2(s23.z, s23.w));
    mx = imax(imax(mx, v.x), imax(v.y, imax(v.z, v.w)));
    mn = imin(imin(mn, v.x), imin(v.y, imin(v.z, v.w)));
    r4[q] = v;
  }
}

// Modulation of the lane's slot, in place; the row holds the block-shifted inputs.  Each half FFT runs in registers and
// is parked over the half of the row it consumed; post_fold() finishes from there.  The loop over the halves is rolled
// (one copy of the FFT code), only the pre-twiddle exists per half.
template <bool SAT>
XB_HD void slot_modulate(i32 *row, const SynTw &tw, const FoldK &fk, const i32 z) {
  int4 *r4 = reinterpret_cast<int4 *>(row);
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
    int4 *h4 = r4 + 16 * half;
    int2 E[32];
    {
      i32 x[64];
#pragma unroll
      for (int q = 0; q < 16; q++) {
        const int4 v = h4[q];
        x[4 * q + 0] = v.x, x[4 * q + 1] = v.y, x[4 * q + 2] = v.z, x[4 * q + 3] = v.w;
      }
      if (half == 0)
        pre_twiddle<SAT, 0>(x, E, tw, z);
      else
        pre_twiddle<SAT, 1>(x, E, tw, z);
    }
    fft_stages<SAT>(E, tw, z);
#pragma unroll
    for (int q = 0; q < 16; q++) h4[q] = make_int4(E[2 * q].x, E[2 * q].y, E[2 * q + 1].x, E[2 * q + 1].y);
  }
  post_fold<SAT>(r4, tw, fk, z);
}

template <int P>
XB_HD void window_unit(const int4 *rows4, int lane, const WinCoef &wc, int16_t *pcm, int ch_fac) {
#pragma unroll 1
  for (int g = 0; g < 8; g++) {
    i32 o[4][2];
    window_group<P, 4>(rows4, chunk_pos(lane), 4 * g, wc, o);
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int slot = 4 * g + t;
      if (ch_fac == 1) {
        *reinterpret_cast<i32 *>(pcm + 64 * slot + 2 * lane) = (o[t][0] & 0xffff) | (i32)((u32)o[t][1] << 16);
      } else {
        pcm[ch_fac * (64 * slot + 2 * lane)] = (int16_t)o[t][0];
        pcm[ch_fac * (64 * slot + 2 * lane + 1)] = (int16_t)o[t][1];
      }
    }
  }
}

}  // namespace syn
}  // namespace xb
