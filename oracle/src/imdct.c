/*
 * oracle/src/imdct.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain-C restatement of the AAC-LC 1024/128 IMDCT + window/overlap-add path of libxaac
 * (SURVEY.md §8a-A).  Index walks of the reference are restated in closed form; every function cites
 * the reference lines it follows (paths relative to /root/reference).  Pinned against the compiled
 * reference (oracle/_ref/libxaac_ref.so : ref_imdct_process) by tests/test_oracle_imdct.py — the
 * reference ships no golden vectors for this path.
 */
#include <string.h>
#include "fixmath.h"
#include "xaac_oracle.h"

/* The first 7500 bytes of ia_aac_dec_imdct_tables_struct (decoder/ixheaacd_aac_rom.h:112-121). */
typedef struct {
  const i16 *cs;        /* cosine_array_2048_256[514] : pairs (A_p, B_p), p = 0..256            */
  const int8_t *dr_long;  /* dig_rev_table8_long[64]                                             */
  const int8_t *dr_short; /* dig_rev_table8_short[8]                                             */
  const i32 *tw;        /* fft_twiddle[448] : lo16 multiplies same component, hi16 the cross one */
  const i16 *win_long[2];  /* [0] sine, [1] KBD (decoder/ixheaacd_aacdecoder.c:192-200)          */
  const i16 *win_short[2];
} imdct_rom;

static void rom_bind(imdct_rom *r, const uint8_t *blob) {
  r->cs = (const i16 *)(blob + XO_ROM_COS);
  r->dr_long = (const int8_t *)(blob + XO_ROM_DIGREV_LONG);
  r->dr_short = (const int8_t *)(blob + XO_ROM_DIGREV_SHORT);
  r->tw = (const i32 *)(blob + XO_ROM_FFT_TW);
  r->win_long[0] = (const i16 *)(blob + XO_ROM_WIN_LONG_SINE);
  r->win_long[1] = (const i16 *)(blob + XO_ROM_WIN_LONG_KBD);
  r->win_short[0] = (const i16 *)(blob + XO_ROM_WIN_SHORT_SINE);
  r->win_short[1] = (const i16 *)(blob + XO_ROM_WIN_SHORT_KBD);
}

/* decoder/ixheaacd_aac_tns.c:422-448 */
int xo_calc_max_spectral_line(const i32 *x, int n) {
  i32 acc = 0;
  for (int i = 0; i < n; i++) acc |= ox_abs_nrm(x[i]);
  return ox_norm32(acc);
}

/* (C,S) pair used by both twiddle passes for complex bin c of an n-point (n = 1024|128) transform.
 * The reference walks the table linearly and swaps the roles of the two halves of each pair after
 * the first entry and again for the last one (aac_imdct.c:180-181 vs :204-205 vs :238-239). */
static void cs_pair(const imdct_rom *r, int n, int c, i16 *C, i16 *S) {
  int st = (n == 1024) ? 2 : 16;
  int q = n >> 2;
  if (c <= q) {
    *C = r->cs[st * c];
    *S = r->cs[st * c + 1];
  } else {
    int p = (n >> 1) - c;
    *C = r->cs[st * p + 1];
    *S = r->cs[st * p];
  }
}

/* decoder/ixheaacd_aac_imdct.c:165-329 — fold + pre-twiddle into n/2 complex bins, block shift by expo */
static void pretwiddle(const imdct_rom *r, const i32 *spec, i32 *out, int n, int expo) {
  for (int c = 0; c < (n >> 1); c++) {
    i16 C, S;
    cs_pair(r, n, c, &C, &S);
    i32 xr = spec[2 * c], xi = spec[n - 1 - 2 * c];
    i32 re = ox_add(ox_mul32x16(xr, C), ox_mul32x16(xi, S));
    i32 im = ox_sub(ox_mul32x16(xi, C), ox_mul32x16(xr, S));
    if (expo < 0) {
      re = ox_shl32(re, -expo);
      im = ox_shl32(im, -expo);
    } else {
      re = ox_shr32(re, expo);
      im = ox_shr32(im, expo);
    }
    out[2 * c] = re;
    out[2 * c + 1] = im;
  }
}

/* Radix-8 butterfly core shared by all stages (aac_imdct.c:876-999, 1029-1150, 1213-1375).
 * late != 0 selects the scaling used after a twiddle multiply, where x1,x2,x4,x6 arrive pre-doubled
 * and x3,x5,x7 do not (aac_imdct.c:1289-1319). Outputs are in storage order base + q*del. */
static void bfly8(i32 *xr, i32 *xi, int late, i32 *yr, i32 *yi) {
#define SH(v, s) ox_lsl((v), (s))
  i32 t;
  /* even half */
  xr[0] = ox_add(xr[0], xr[4]); xi[0] = ox_add(xi[0], xi[4]);
  xr[4] = ox_sub(xr[0], SH(xr[4], 1)); xi[4] = ox_sub(xi[0], SH(xi[4], 1));
  xr[2] = ox_add(xr[2], xr[6]); xi[2] = ox_add(xi[2], xi[6]);
  xr[6] = ox_sub(xr[2], SH(xr[6], 1)); xi[6] = ox_sub(xi[2], SH(xi[6], 1));
  xr[0] = ox_add(xr[0], xr[2]); xi[0] = ox_add(xi[0], xi[2]);
  xr[2] = ox_sub(xr[0], SH(xr[2], 1)); xi[2] = ox_sub(xi[0], SH(xi[2], 1));
  xr[4] = ox_add(xr[4], xi[6]); xi[4] = ox_sub(xi[4], xr[6]);
  t = xr[6];
  xr[6] = ox_sub(xr[4], SH(xi[6], 1)); xi[6] = ox_add(xi[4], SH(t, 1));
  /* odd half */
  int a = late ? 1 : 0; /* extra doubling applied to the un-doubled operands */
  xr[1] = ox_add(xr[1], SH(xr[5], a)); xi[1] = ox_add(xi[1], SH(xi[5], a));
  xr[5] = ox_sub(xr[1], SH(xr[5], a + 1)); xi[5] = ox_sub(xi[1], SH(xi[5], a + 1));
  xr[3] = ox_add(xr[3], xr[7]); xi[3] = ox_add(xi[3], xi[7]);
  xr[7] = ox_sub(xr[3], SH(xr[7], 1)); xi[7] = ox_sub(xi[3], SH(xi[7], 1));
  xr[1] = ox_add(xr[1], SH(xr[3], a)); xi[1] = ox_add(xi[1], SH(xi[3], a));
  xr[3] = ox_sub(xr[1], SH(xr[3], a + 1)); xi[3] = ox_sub(xi[1], SH(xi[3], a + 1));
  xr[5] = ox_add(xr[5], xi[5]); xi[5] = ox_sub(xr[5], SH(xi[5], 1));
  xr[7] = ox_add(xr[7], xi[7]); xi[7] = ox_sub(xr[7], SH(xi[7], 1));
  xi[7] = ox_sub(xr[5], SH(xi[7], a)); xr[5] = ox_sub(xi[7], SH(xr[5], 1));
  xi[5] = ox_sub(SH(xr[7], a), xi[5]); xr[7] = ox_sub(xi[5], SH(xr[7], a + 1));
  xi[7] = SH(xi[7], 1); xr[5] = SH(xr[5], 1); xi[5] = SH(xi[5], 1); xr[7] = SH(xr[7], 1);
  /* combine */
  xr[0] = ox_add(xr[0], xr[1]); xi[0] = ox_add(xi[0], xi[1]);
  xr[1] = ox_sub(xr[0], SH(xr[1], 1)); xi[1] = ox_sub(xi[0], SH(xi[1], 1));
  xr[2] = ox_add(xr[2], xi[3]);
  t = ox_sub(xr[2], SH(xi[3], 1));
  xi[2] = ox_sub(xi[2], xr[3]);
  xi[3] = ox_add(xi[2], SH(xr[3], 1));
  yr[0] = xr[0]; yi[0] = xi[0];
  yr[2] = xr[2]; yi[2] = xi[2];
  yr[4] = xr[1]; yi[4] = xi[1];
  yr[6] = t;     yi[6] = xi[3];
  const i32 k = 0x5A82;
  xi[7] = ox_add(xr[4], ox_mul32x16l(xi[7], k)); xr[4] = ox_sub(xi[7], SH(xr[4], 1));
  xr[7] = ox_add(xi[4], ox_mul32x16l(xr[7], k)); xi[4] = ox_sub(xr[7], SH(xi[4], 1));
  xi[5] = ox_add(xr[6], ox_mul32x16l(xi[5], k)); xr[6] = ox_sub(xi[5], SH(xr[6], 1));
  xr[5] = ox_add(xi[6], ox_mul32x16l(xr[5], k)); xi[6] = ox_sub(xr[5], SH(xi[6], 1));
  yr[1] = xi[7]; yi[1] = xr[7];
  yr[3] = xi[5]; yi[3] = xr[5];
  yr[5] = ox_sub(0, xr[4]); yi[5] = ox_sub(0, xi[4]);
  yr[7] = ox_sub(0, xr[6]); yi[7] = ox_sub(0, xi[6]);
#undef SH
}

/* twiddle multiply of aac_imdct.c:1179-1185 (pre-doubled) and :1256-1260 (plain) */
static void tw_mul(i32 *re, i32 *im, i32 w, int dbl) {
  i32 a = ox_sub(ox_mul32x16l(*re, w), ox_mul32x16h(*im, w));
  i32 b = ox_add(ox_mul32x16h(*re, w), ox_mul32x16l(*im, w));
  *re = dbl ? ox_shl1(a) : a;
  *im = dbl ? ox_shl1(b) : b;
}

/* decoder/ixheaacd_aac_imdct.c:834-1622 — np = 512 (three radix-8 stages) or 64 (two) */
static void fft_r8(const imdct_rom *r, int np, const i32 *x, i32 *y) {
  const int8_t *dr = (np == 512) ? r->dr_long : r->dr_short;
  i32 xr[8], xi[8], yr[8], yi[8];
  /* stage 1: digit-reversed gather, natural-order scatter (:856-1000) */
  for (int g = 0; g < np / 8; g++) {
    int b = dr[g];
    for (int q = 0; q < 8; q++) {
      /* x0,x2,x4,x6 at b + {0,1,2,3}*np/4 ; x1,x3,x5,x7 at b + np/8 + {0,1,2,3}*np/4 */
      int idx = b + (q >> 1) * (np >> 2) + (q & 1) * (np >> 3);
      xr[q] = x[2 * idx];
      xi[q] = x[2 * idx + 1];
    }
    bfly8(xr, xi, 0, yr, yi);
    for (int q = 0; q < 8; q++) {
      y[2 * (8 * g + q)] = yr[q];
      y[2 * (8 * g + q) + 1] = yi[q];
    }
  }
  /* later stages, in place (:1007-1384 middle, :1386-1621 last). Twiddle for leg q of column m is
   * tw[q*m*(64/del)]; column 0 of a non-final stage skips the multiply (:1011-1152). */
  for (int del = 8; del < np; del <<= 3) {
    int last = (del * 8 == np);
    int step = 64 / del; /* 8 for del=8, 1 for del=64 */
    for (int m = 0; m < del; m++) {
      int plain = (m == 0 && !last);
      for (int base = m; base < np; base += 8 * del) {
        for (int q = 0; q < 8; q++) {
          int idx = base + q * del;
          xr[q] = y[2 * idx];
          xi[q] = y[2 * idx + 1];
          if (!plain && q > 0) tw_mul(&xr[q], &xi[q], r->tw[q * m * step], (q == 1) || !(q & 1));
        }
        bfly8(xr, xi, !plain, yr, yi);
        for (int q = 0; q < 8; q++) {
          int idx = base + q * del;
          y[2 * idx] = yr[q];
          y[2 * idx + 1] = yi[q];
        }
      }
    }
  }
}

/* decoder/ixheaacd_aac_imdct.c:1657-1670 : spec -> (scratch) -> spec ; returns expo+2 */
int xo_inverse_transform(const uint8_t *rom, i32 *spec, i32 *scratch, int expo, int n) {
  imdct_rom r;
  rom_bind(&r, rom);
  pretwiddle(&r, spec, scratch, n, expo);
  fft_r8(&r, n >> 1, scratch, spec);
  return expo + 2;
}

/* decoder/ixheaacd_aac_imdct.c:331-504 — n = 1024 (adjust 50) or 128 (adjust 402) */
static void post_twiddle(const imdct_rom *r, i32 *out, const i32 *y, int n) {
  i16 adj = (n == 1024) ? 50 : 402;
  for (int c = 0; c < (n >> 1); c++) {
    i16 C, S;
    cs_pair(r, n, c, &C, &S);
    i32 yr = y[2 * c], yi = y[2 * c + 1];
    i32 orr = ox_add(ox_mul32x16(yr, C), ox_mul32x16(yi, S));
    i32 oi = ox_sub(ox_mul32x16(yr, S), ox_mul32x16(yi, C));
    i32 t1 = ox_mul32x16(oi, (i16)-adj), t2 = ox_mul32x16(orr, adj);
    out[2 * c] = ox_add(orr, t1);
    out[n - 1 - 2 * c] = ox_add(oi, t2);
  }
}
void xo_post_twiddle(const uint8_t *rom, i32 *out, const i32 *y, int n) {
  imdct_rom r;
  rom_bind(&r, rom);
  post_twiddle(&r, out, y, n);
}

/* decoder/ixheaacd_lpfuncs.c:316-323 */
static void spec_to_overlap(i32 *ovl, const i32 *x, int q_shift, int n) {
  for (int i = 0; i < n; i++) ovl[i] = ox_shr32_sat(x[i], 16 - q_shift);
}

/* decoder/ixheaacd_aac_imdct.c:506-832 — long->long fused post-twiddle + window + OLA.
 * t[] is the post-twiddled block (the reference never materialises it). */
static void long_long_ola(const i32 *t, i32 *ovl, i32 *out, const i16 *win, int q_shift, int ch_fac) {
  for (int m = 0; m < 512; m++) {
    i32 x = t[512 + m];
    i16 wlo = win[2 * m], whi = win[2 * m + 1];
    i32 prev = ovl[511 - m];
    i32 a, b;
    if (q_shift > 0) {
      a = ox_shl32_sat(ox_mul32x16(x, wlo), q_shift);
      b = ox_shl32_sat(ox_mul32x16(ox_neg_sat(x), whi), q_shift);
    } else {
      prev = (i16)prev; /* aac_imdct.c:679 — overlap read through a WORD16 in this branch */
      a = ox_shr32(ox_mul32x16(x, wlo), -q_shift);
      b = ox_shr32(ox_mul32x16(ox_neg_sat(x), whi), -q_shift);
    }
    out[ch_fac * m] = ox_sub_sat(a, ox_mul32x16_fullsat(prev, whi));
    out[ch_fac * (1023 - m)] = ox_sub_sat(b, ox_mul32x16_fullsat(prev, wlo));
  }
  for (int k = 0; k < 512; k++) ovl[k] = ox_shr32_sat(t[k], 16 - q_shift);
}

/* decoder/ixheaacd_block.c:1193-1218 */
static void ola1(const i32 *coef, const i32 *prev, i32 *out, const i16 *w, int q_shift, int size, int ch_fac) {
  for (int i = 0; i < size; i++) {
    i16 w1 = w[2 * size - 2 * i - 1], w2 = w[2 * size - 2 * i - 2];
    i32 c = coef[2 * size - 1 - i];
    out[ch_fac * (size - 1 - i)] = ox_sub_sat(ox_shl32_dir_sat_limit(ox_mul32x16(c, w2), q_shift),
                                              ox_add_sat(0, ox_mul32x16_fullsat(prev[i], w1)));
    out[ch_fac * (size + i)] = ox_sub_sat(ox_shl32_dir_sat_limit(ox_mul32x16(ox_neg_sat(c), w1), q_shift),
                                          ox_add_sat(0, ox_mul32x16_fullsat(prev[i], w2)));
  }
}

/* decoder/ixheaacd_block.c:1220-1240 */
static void ola2(const i32 *coef, const i32 *prev, i32 *out, const i16 *w, int q_shift, int size) {
  for (int i = 0; i < size; i++) {
    i32 acc = ox_sub_sat(ox_mul32x16(coef[size + i], w[2 * i]), ox_mul32x16(prev[size - 1 - i], w[2 * i + 1]));
    out[i] = ox_shr32_sat(acc, 16 - (q_shift + 1));
  }
  for (int i = 0; i < size; i++) {
    i32 acc = ox_sub_sat(ox_mul32x16(ox_neg_sat(coef[2 * size - 1 - i]), w[2 * size - 2 * i - 1]),
                         ox_mul32x16(prev[i], w[2 * size - 2 * i - 2]));
    out[i + size] = ox_shr32_sat(acc, 16 - (q_shift + 1));
  }
}

/* decoder/ixheaacd_lpfuncs.c:94-178 (size_01 = 64) */
static void process_win_seq(const i32 *coef, const i32 *prev, i32 *out, const i16 *wl, const i16 *ws,
                            int q_shift, int ch_fac, int flag) {
  const int s1 = 64, s7 = 448, s8 = 512, s9 = 576, s14 = 896, s15 = 960;
  const i16 *w_sh, *w_lg;
  const i32 *pv;
  if (flag) {
    for (int i = 0; i < s7; i++) {
      i32 t = ox_shl32_dir_sat_limit(ox_mul32x16(coef[s8 + i], wl[2 * i]), q_shift + 1);
      out[ch_fac * i] = ox_add_sat(t, ox_lsl(prev[i], 16));
      i32 a = ox_shl32_dir_sat_limit(ox_mul32x16(ox_sub(0, coef[s15 - 1 - i]), wl[2 * (s7 - i) - 1]), q_shift);
      out[ch_fac * (i + s9)] = ox_shl1(a);
    }
    w_sh = ws;
    w_lg = wl + s14;
    pv = prev + s8 - 1;
  } else {
    for (int i = 0; i < s7; i++) {
      out[ch_fac * i] = ox_mul32x16_fullsat(prev[s8 - 1 - i], ox_neg16(wl[2 * i + 1]));
      out[ch_fac * (s9 + i)] = ox_sub_sat(ox_shl32_dir_sat_limit(ox_sub(0, coef[s15 - 1 - i]), q_shift - 1),
                                          ox_mul32x16_fullsat(prev[i + s1], wl[2 * s7 - 2 - 2 * i]));
    }
    w_sh = wl + s14;
    w_lg = ws;
    pv = prev + s1 - 1;
  }
  for (int k = 0; k < s1; k++) {
    i32 c = coef[s15 + k];
    i16 win1 = w_lg[2 * k], win2 = w_lg[2 * k + 1];
    i16 win4 = w_sh[2 * k], win3 = w_sh[2 * k + 1];
    i32 p = pv[-k];
    i32 a = ox_sub_sat(ox_shl32_dir_sat_limit(ox_mul32x16(c, win1), q_shift), ox_mul32x16_fullsat(p, win3));
    out[ch_fac * (s7 + k)] = ox_lsl(a, flag);
    a = ox_sub_sat(ox_shl32_dir_sat_limit(ox_mul32x16(ox_neg_sat(c), win2), q_shift), ox_mul32x16_fullsat(p, win4));
    out[ch_fac * (s9 - 1 - k)] = ox_lsl(a, flag);
  }
}

/* decoder/ixheaacd_lpfuncs.c:180-216 */
static void long_short_win_process(const i32 *cur, const i32 *prev, i32 *out, const i16 *sw, const i16 *lwp,
                                   int q_shift, int ch_fac, int flag) {
  const int s1 = 64, s2 = 128, s3 = 192;
  for (int i = s1 - 1; i >= 0; i--) {
    i32 c1 = cur[s3 - 1 - (s1 - 1 - i)];
    i32 c2 = cur[-s1 + (s1 - 1 - i)];
    i16 sh1 = sw[s2 - 1 - 2 * (s1 - 1 - i)];
    i16 sh2 = sw[s2 - 2 - 2 * (s1 - 1 - i)];
    i32 a = ox_sub_sat(
        ox_shl32_dir_sat_limit(ox_sub(ox_mul32x16(c1, sh2), ox_mul32x16(c2, sh1)), q_shift),
        ox_mul32x16_fullsat(prev[i], lwp[0 - 2 - 2 * i]));
    out[ch_fac * i] = a;
    if (flag) {
      a = ox_sub_sat(
          ox_shl32_dir_sat_limit(ox_sub(ox_mul32x16(ox_neg_sat(c1), sh1), ox_mul32x16(c2, sh2)), q_shift),
          ox_mul32x16_fullsat(prev[s2 - 1 - i], lwp[-2 * s2 + 2 * i]));
      out[ch_fac * (s2 - 1 - i)] = a;
    }
  }
}

/* decoder/ixheaacd_lpfuncs.c:218-284 */
static void long_short_win_seq(const i32 *cur, i32 *prev, i32 *out, const i16 *sw, const i16 *swp,
                               const i16 *lwp, int q_shift, int ch_fac) {
  const int s1 = 64, s2 = 128, s6 = 384, s7 = 448, s8 = 512, s9 = 576, s10 = 640, s16 = 1024;
  for (int i = 0; i < s7; i++) out[ch_fac * i] = ox_mul32x16_fullsat(prev[s8 - 1 - i], ox_neg16(lwp[2 * i + 1]));
  for (int i = 0; i < s1; i++)
    out[ch_fac * (s7 + i)] =
        ox_sub_sat(ox_shl32_dir_sat_limit(ox_mul32x16(cur[s1 + i], swp[2 * i]), q_shift),
                   ox_mul32x16_fullsat(prev[s1 - 1 - i], lwp[2 * s7 + 1 + 2 * i]));
  for (int i = 0; i < s1; i++)
    out[ch_fac * (s8 + i)] =
        ox_sub_sat(ox_shl32_dir_sat_limit(ox_mul32x16(ox_neg_sat(cur[s2 - 1 - i]), swp[s2 - 2 * i - 1]), q_shift),
                   ox_mul32x16_fullsat(prev[i], lwp[s16 - 2 - 2 * i]));
  for (int b = 0; b < 4; b++) {
    int inc = b * s2;
    long_short_win_process(cur + s1 + inc, prev + s1 + inc, out + ch_fac * (s9 + inc), sw,
                           lwp + 2 * (s7 - inc), q_shift, ch_fac, b != 3);
  }
  for (int i = 0; i < s1; i++) {
    i32 a = ox_sub(ox_mul32x16(ox_sub(0, cur[s10 - 1 - i]), sw[s2 - 2 * i - 1]),
                   ox_mul32x16(cur[s6 + i], sw[s2 - 2 * i - 2]));
    prev[i] = ox_round16(ox_shl32_dir_sat_limit(a, q_shift + 1));
  }
}

/* decoder/ixheaacd_lpfuncs.c:347-802, 1024-sample frames only (AAC-LC / HE-AAC core). */
int xo_imdct_process(const uint8_t *rom, i32 *spec, i32 *ovl, i32 *prev_shape, i32 *prev_seq, int win_seq,
                     int win_shape, i32 *out, int ch_fac) {
  imdct_rom r;
  rom_bind(&r, rom);
  i32 scratch[1024];
  const i16 *wl = r.win_long[*prev_shape];
  const i16 *wsp = r.win_short[*prev_shape];
  const int s1 = 64, s2 = 128, s6 = 384, s7 = 448, s8 = 512, s9 = 576, s10 = 640, s14 = 896, s15 = 960;
  int pseq = *prev_seq;
  int prev_longish = (pseq == XO_ONLY_LONG || pseq == XO_LONG_STOP);
  int adj = 0;

  if (win_seq != XO_EIGHT_SHORT) {
    int expo = 8 - (xo_calc_max_spectral_line(spec, 1024) - 1);
    int imdct_scale = xo_inverse_transform(rom, spec, scratch, expo, 1024);
    int q_shift = (31 + imdct_scale) - 26;
    switch (win_seq) {
      case XO_ONLY_LONG:
        if (prev_longish) {
          post_twiddle(&r, scratch, spec, 1024); /* fused in the reference (aac_imdct.c:506) */
          long_long_ola(scratch, ovl, out, wl, q_shift, ch_fac);
          adj = 2;
        } else {
          post_twiddle(&r, scratch, spec, 1024);
          process_win_seq(scratch, ovl, out, wl, wsp, q_shift, ch_fac, 1);
          spec_to_overlap(ovl, scratch, q_shift, s8);
          adj = 1;
        }
        break;
      case XO_LONG_START:
        post_twiddle(&r, scratch, spec, 1024);
        if (prev_longish) {
          ola1(scratch, ovl, out, wl, q_shift, s8, ch_fac);
          adj = 2;
        } else {
          process_win_seq(scratch, ovl, out, wl, wsp, q_shift, ch_fac, 1);
          adj = 1;
        }
        /* lpfuncs.c:286-295 (nolap1_32) + :571 */
        for (int i = 0; i < s7; i++) ovl[i] = ox_shr32_sat(ox_neg_sat(scratch[s1 + s7 - 1 - i]), 16 - q_shift);
        spec_to_overlap(ovl + s7, scratch, q_shift, s1);
        break;
      case XO_LONG_STOP:
        post_twiddle(&r, scratch, spec, 1024);
        if (!prev_longish) {
          for (int i = 0; i < s7; i++) out[ch_fac * i] = ox_shl32_sat((i16)ovl[i], 15); /* lpfuncs.c:325-333 */
          ola1(scratch + s14, ovl + s7, out + ch_fac * s7, wsp, q_shift, s1, ch_fac);
          for (int i = 0; i < s7; i++) /* lpfuncs.c:297-304 */
            out[ch_fac * (s9 + i)] = ox_shl32_dir_sat_limit(ox_neg_sat(scratch[s8 + s7 - 1 - i]), q_shift - 1);
        } else {
          process_win_seq(scratch, ovl, out, wl, wsp, q_shift, ch_fac, 0);
        }
        adj = 2;
        spec_to_overlap(ovl, scratch, q_shift, s8);
        break;
    }
  } else {
    const i16 *sw = r.win_short[win_shape];
    int expo = 5 - (xo_calc_max_spectral_line(spec, 1024) - 1);
    int scale0 = 0;
    for (int w = 0; w < 8; w++) {
      int sc = xo_inverse_transform(rom, spec + w * s2, scratch + w * s2, expo, 128);
      if (w == 0) scale0 = sc;
      post_twiddle(&r, scratch + w * s2, spec + w * s2, 128);
    }
    int q_shift = 31 + scale0 - 23;
    if (!prev_longish) {
      i32 loc[64];
      for (int i = 0; i < s7; i++) out[ch_fac * i] = ox_shl32_sat((i16)ovl[i], 15);
      ola1(scratch, ovl + s7, out + ch_fac * s7, wsp, q_shift, s1, ch_fac);
      for (int b = 0; b < 3; b++) {
        int inc = b * s2;
        spec_to_overlap(loc, scratch + inc, q_shift, s1);
        ola1(scratch + s2 + inc, loc, out + ch_fac * (s9 + inc), sw, q_shift, s1, ch_fac);
      }
      ola2(scratch + s8, scratch + s6, ovl, sw, q_shift, s1);
      for (int i = 0; i < s1; i++) { /* lpfuncs.c:335-345 */
        out[ch_fac * (s15 + i)] = ox_shl32_sat((i16)ovl[i], 15);
        ovl[i] = ovl[s1 + i];
      }
    } else {
      long_short_win_seq(scratch, ovl, out, sw, wsp, wl, q_shift, ch_fac);
    }
    adj = 2;
    for (int b = 0; b < 3; b++) {
      int inc = b * s2;
      ola2(scratch + s10 + inc, scratch + s8 + inc, ovl + s1 + inc, sw, q_shift, s1);
    }
    spec_to_overlap(ovl + s7, scratch + s14, q_shift, s1);
  }
  *prev_shape = win_shape;
  *prev_seq = win_seq;
  return adj;
}

void xo_imdct_process_batch(const uint8_t *rom, i32 *spec, i32 *ovl, i32 *prev_shape, i32 *prev_seq,
                            const i32 *win_seq, const i32 *win_shape, i32 *out, i32 *qshift_adj, int n) {
  for (int u = 0; u < n; u++)
    qshift_adj[u] = xo_imdct_process(rom, spec + (size_t)u * 1024, ovl + (size_t)u * 512, prev_shape + u,
                                     prev_seq + u, win_seq[u], win_shape[u], out + (size_t)u * 1024, 1);
}
