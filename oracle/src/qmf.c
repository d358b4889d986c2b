/*
 * oracle/src/qmf.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain-C restatement of the fixed-point (WORD32/WORD16) complex "HQ" SBR QMF banks of libxaac
 * (SURVEY.md §8a-C): 32-band analysis and 64-band synthesis with their cosine/sine modulation.
 * Pointer walks of the reference are restated with explicit indices; every function cites the reference
 * lines it follows (paths relative to /root/reference). Pinned against the compiled reference
 * (oracle/_ref : ref_cos_sin_mod / ref_synt_qmffilt_hq / ref_anal_qmffilt_hq) by tests/test_oracle_qmf.py.
 */
#include <string.h>
#include "fixmath.h"
#include "xaac_oracle.h"

#define T16(off) ((const i16 *)(qrom + (off)))
#define T32(off) ((const i32 *)(qrom + (off)))

/* decoder/generic/ixheaacd_qmf_dec_generic.c:1736-1829 — in-place radix-4 stage on interleaved complex x.
 * `groups` blocks of 4*span points; leg distance = span; twiddles (si,co) x3 per butterfly position. */
static void radix4_stage(const i16 *w, i32 *x, int groups, int span) {
  for (int g = 0; g < groups; g++) {
    for (int i = 0; i < span; i++) {
      i32 *e0 = x + 2 * (g * 4 * span + i), *e1 = e0 + 2 * span, *e2 = e0 + 4 * span, *e3 = e0 + 6 * span;
      const i16 *tw = w + 6 * i;
      i16 si1 = tw[0], co1 = tw[1], si2 = tw[2], co2 = tw[3], si3 = tw[4], co3 = tw[5];
      i32 xh0 = ox_add_sat(e0[0], e2[0]), xl0 = ox_sub_sat(e0[0], e2[0]);
      i32 xh20 = ox_add_sat(e1[0], e3[0]), xl20 = ox_sub_sat(e1[0], e3[0]);
      i32 xh1 = ox_add_sat(e0[1], e2[1]), xl1 = ox_sub_sat(e0[1], e2[1]);
      i32 xh21 = ox_add_sat(e1[1], e3[1]), xl21 = ox_sub_sat(e1[1], e3[1]);
      i32 xt0 = ox_sub_sat(xh0, xh20), yt0 = ox_sub_sat(xh1, xh21);
      i32 xt1 = ox_add_sat(xl0, xl21), xt2 = ox_sub_sat(xl0, xl21);
      i32 yt2 = ox_add_sat(xl1, xl20), yt1 = ox_sub_sat(xl1, xl20);
      e0[0] = ox_add_sat(xh0, xh20);
      e0[1] = ox_add_sat(xh1, xh21);
      e3[0] = ox_shl1(ox_add(ox_mul32x16(yt2, si3), ox_mul32x16(xt2, co3)));
      e3[1] = ox_shl1(ox_sub(ox_mul32x16(yt2, co3), ox_mul32x16(xt2, si3)));
      e2[0] = ox_shl1(ox_add(ox_mul32x16(yt0, si2), ox_mul32x16(xt0, co2)));
      e2[1] = ox_shl1(ox_sub(ox_mul32x16(yt0, co2), ox_mul32x16(xt0, si2)));
      e1[0] = ox_shl1(ox_add(ox_mul32x16(yt1, si1), ox_mul32x16(xt1, co1)));
      e1[1] = ox_shl1(ox_sub(ox_mul32x16(yt1, co1), ox_mul32x16(xt1, si1)));
    }
  }
}

/* generic:1934-2015 — final radix-2 stage of the 32-point FFT with digit-reversed scatter */
static void post_radix2_32(i32 *y, const i32 *x, const i32 *digrev) {
  i32 *y0 = y, *y1 = y + 8, *y2 = y + 32, *y3 = y + 40;
  for (int blk = 0; blk < 4; blk++) { /* (k,i) = (0,0),(0,8),(1,0),(1,8) */
    int h2 = digrev[blk] >> 2;
    const i32 *a = x + (blk >> 1) * 32 + (blk & 1) * 8; /* x0 walk */
    const i32 *b = a + 16;                               /* x2 walk */
    for (int half = 0; half < 2; half++) {
      const i32 *c = half ? b : a;
      int o = h2 + 2 * half;
      y0[o] = ox_add_sat(c[0], c[2]); y0[o + 1] = ox_add_sat(c[1], c[3]);
      y2[o] = ox_sub_sat(c[0], c[2]); y2[o + 1] = ox_sub_sat(c[1], c[3]);
      y1[o] = ox_add_sat(c[4], c[6]); y1[o + 1] = ox_add_sat(c[5], c[7]);
      y3[o] = ox_sub_sat(c[4], c[6]); y3[o + 1] = ox_sub_sat(c[5], c[7]);
    }
  }
}

/* generic:1831-1932 — final radix-4 stage (no twiddles) of the 16-point FFT with digit-reversed scatter */
static void post_radix4_16(i32 *y, const i32 *x, const i32 *digrev) {
  i32 *y0 = y, *y1 = y + 8, *y2 = y + 16, *y3 = y + 24;
  for (int k = 0; k < 2; k++) {
    int h2 = digrev[k] >> 2;
    for (int half = 0; half < 2; half++) {
      const i32 *c = x + 16 * k + 8 * half;
      int o = h2 + 2 * half;
      i32 xh0 = ox_add_sat(c[0], c[4]), xh1 = ox_add_sat(c[1], c[5]);
      i32 xl0 = ox_sub_sat(c[0], c[4]), xl1 = ox_sub_sat(c[1], c[5]);
      i32 zh0 = ox_add_sat(c[2], c[6]), zh1 = ox_add_sat(c[3], c[7]);
      i32 zl0 = ox_sub_sat(c[2], c[6]), zl1 = ox_sub_sat(c[3], c[7]);
      y0[o] = ox_add_sat(xh0, zh0); y0[o + 1] = ox_add_sat(xh1, zh1);
      y1[o] = ox_add_sat(xl0, zl1); y1[o + 1] = ox_sub_sat(xl1, zl0);
      y2[o] = ox_sub_sat(xh0, zh0); y2[o + 1] = ox_sub_sat(xh1, zh1);
      y3[o] = ox_sub_sat(xl0, zl1); y3[o + 1] = ox_add_sat(xl1, zl0);
    }
  }
}

/* generic:259-466 — complex exponential modulation of one slot: subband[0..2M-1] (first half) and
 * subband[64..64+2M-1] (second half), M = no_channels/2 = 32 (synthesis) or 16 (analysis). */
void xo_cos_sin_mod(const uint8_t *qrom, i32 *sb, int no_channels) {
  const int M = no_channels >> 1, N = 2 * M;
  const i16 *tw = T16(no_channels == 64 ? XO_QROM_SINCOS_L64 : XO_QROM_SINCOS_L32);
  const i16 *alt = T16(no_channels == 64 ? XO_QROM_ALTSIN_L64 : XO_QROM_ALTSIN_L32);
  i32 t[128];
  i32 *s1 = sb, *s2 = sb + 64, *t1 = t, *t2 = t + 64;
  /* pre-twiddle (:290-367): step n pairs sample n with sample N-1-n; even steps fill T from the front,
   * odd steps from the back */
  for (int n = 0; n < M; n++) {
    i16 wim = tw[2 * n], wre = tw[2 * n + 1];
    i32 a = s1[n], b = s1[N - 1 - n], c = s2[n], d = s2[N - 1 - n];
    if (!(n & 1)) {
      int j = n >> 1;
      t1[2 * j] = ox_add_sat(ox_mul32x16(a, wre), ox_mul32x16(b, wim));
      t1[2 * j + 1] = ox_sub_sat(ox_mul32x16(b, wre), ox_mul32x16(a, wim));
      t2[2 * j] = ox_sub_sat(ox_mul32x16(d, wim), ox_mul32x16(c, wre));
      t2[2 * j + 1] = ox_add_sat(ox_mul32x16(c, wim), ox_mul32x16(d, wre));
    } else {
      int j = (n - 1) >> 1;
      t1[N - 1 - 2 * j] = ox_sub_sat(ox_mul32x16(a, wre), ox_mul32x16(b, wim));
      t1[N - 2 - 2 * j] = ox_add_sat(ox_mul32x16(b, wre), ox_mul32x16(a, wim));
      t2[N - 1 - 2 * j] = ox_add_sat(ox_mul32x16(d, wim), ox_mul32x16(c, wre));
      t2[N - 2 - 2 * j] = ox_sub_sat(ox_mul32x16(c, wim), ox_mul32x16(d, wre));
    }
  }
  /* M-point complex FFT of each half (:369-386) */
  if (M == 32) {
    const i16 *w = T16(XO_QROM_W32);
    const i32 *dr = T32(XO_QROM_DIGREV2_32);
    for (int h = 0; h < 2; h++) {
      radix4_stage(w, t + 64 * h, 1, 8);
      radix4_stage(w + 48, t + 64 * h, 4, 2);
      post_radix2_32(sb + 64 * h, t + 64 * h, dr);
    }
  } else {
    const i16 *w = T16(XO_QROM_W16);
    const i32 *dr = T32(XO_QROM_DIGREV4_16);
    for (int h = 0; h < 2; h++) {
      radix4_stage(w, t + 64 * h, 1, 4);
      post_radix4_16(sb + 64 * h, t + 64 * h, dr);
    }
  }
  /* post-twiddle (:388-465), restated out of place: every output depends only on the FFT output f */
  i32 f1[64], f2[64];
  memcpy(f1, s1, sizeof(i32) * N);
  memcpy(f2, s2, sizeof(i32) * N);
  const int H = M >> 1; /* M_2 */
  s1[0] = f1[0] >> 1;
  s1[N - 1] = ox_neg_sat(f1[1] >> 1);
  s2[N - 1] = ox_neg_sat(f2[0] >> 1);
  s2[0] = f2[1] >> 1;
  for (int u = 0; u < H; u++) {
    /* back pair (f[N-2-2u], f[N-1-2u]) with alt[u] */
    i16 wim = alt[2 * u], wre = alt[2 * u + 1];
    i32 re = f1[N - 1 - 2 * u], im = f1[N - 2 - 2 * u];
    s1[N - 2 - 2 * u] = ox_add_sat(ox_mul32x16(re, wre), ox_mul32x16(im, wim));
    s1[1 + 2 * u] = ox_sub_sat(ox_mul32x16(im, wre), ox_mul32x16(re, wim));
    re = f2[N - 1 - 2 * u];
    im = f2[N - 2 - 2 * u];
    s2[1 + 2 * u] = ox_neg_sat(ox_add_sat(ox_mul32x16(re, wre), ox_mul32x16(im, wim)));
    s2[N - 2 - 2 * u] = ox_sub_sat(ox_mul32x16(re, wim), ox_mul32x16(im, wre));
    if (u + 1 < H) {
      /* front pair (f[2+2u], f[3+2u]) with the same alt[u] */
      i32 fim = f1[2 + 2 * u], fre = f1[3 + 2 * u];
      s1[2 + 2 * u] = ox_add_sat(ox_mul32x16(fre, wim), ox_mul32x16(fim, wre));
      s1[N - 3 - 2 * u] = ox_sub_sat(ox_mul32x16(fim, wim), ox_mul32x16(fre, wre));
      fim = f2[2 + 2 * u];
      fre = f2[3 + 2 * u];
      s2[N - 3 - 2 * u] = ox_neg_sat(ox_add_sat(ox_mul32x16(fre, wim), ox_mul32x16(fim, wre)));
      s2[2 + 2 * u] = ox_sub_sat(ox_mul32x16(fre, wre), ox_mul32x16(fim, wim));
    }
  }
}

/* decoder/ixheaacd_env_calc.c:1099-1157, complex variant: wrapping left shift / arithmetic right shift */
static void adjust_scale_cplx(i32 *matrix, int b0, int b1, int s0, int s1, int shift) {
  if (shift == 0) return;
  if (shift > 31) shift = 31;
  if (shift < -31) shift = -31;
  for (int l = s0; l < s1; l++)
    for (int k = b0; k < b1; k++) {
      i32 *re = matrix + 128 * l + k, *im = re + 64;
      if (shift > 0) {
        *re = ox_lsl(*re, shift);
        *im = ox_lsl(*im, shift);
      } else {
        *re = *re >> -shift;
        *im = *im >> -shift;
      }
    }
}

/* decoder/ixheaacd_qmf_dec.c:811-1129 — complex 64-band synthesis, no PS/DRC, non-ELD object types */
void xo_synt_qmffilt_hq(const uint8_t *qrom, i32 *matrix, i16 *fs, i32 *drc_offset, i32 *filter_pos, const i32 *sf,
                        int lsb, int usb, int split, i16 *time_out, int ch_fac) {
  const i16 *qmf_c = T16(XO_QROM_QMF_C);
  int ov_lb_scale = sf[0], lb_scale = sf[1], hb_scale = sf[2], st_syn = sf[3];
  int ov_lb_shift = (st_syn - ov_lb_scale) - 8;  /* :924-926 */
  int lb_shift = (st_syn - lb_scale) - 8;
  int hb_shift = (st_syn - hb_scale) - 8;
  int out_shift = -(st_syn - 3) + 1;             /* :914, :1055 */
  if (ov_lb_shift == lb_shift) {
    adjust_scale_cplx(matrix, 0, lsb, 0, 32, ov_lb_shift);
  } else {
    adjust_scale_cplx(matrix, 0, lsb, 0, split, ov_lb_shift);
    adjust_scale_cplx(matrix, 0, lsb, split, 32, lb_shift);
  }
  adjust_scale_cplx(matrix, lsb, usb, 0, 32, hb_shift);
  int off = *drc_offset, fpos = *filter_pos;
  for (int i = 0; i < 32; i++) {
    i32 *re = matrix + 128 * i, *im = re + 64;
    xo_cos_sin_mod(qrom, re, 64);
    /* generic:1638-1670 — fold to 128 WORD16 state samples */
    i16 *st = fs + off;
    for (int j = 0; j < 32; j++) {
      i32 r1 = re[j], i1 = im[j], r2 = re[63 - j], i2 = im[63 - j];
      st[64 + 63 - j] = ox_round16(ox_shl32_sat(ox_add_sat(i1, r1), out_shift));
      st[63 - j] = ox_round16(ox_shl32_sat(ox_sub_sat(i2, r2), out_shift));
      st[j] = ox_round16(ox_shl32_sat(ox_sub_sat(i1, r1), out_shift));
      st[64 + j] = ox_round16(ox_shl32_sat(ox_add_sat(i2, r2), out_shift));
    }
    /* generic:1508-1542 — 10-tap polyphase window, fp1/fp2 alternate between the two 64-sample phases */
    const i16 *fp1 = fs + ((i & 1) ? 64 : 0), *fp2 = fs + ((i & 1) ? 0 : 64);
    const i16 *c = qmf_c + fpos;
    i16 *out = time_out + ch_fac * 64 * i;
    for (int k = 0; k < 64; k++) {
      i32 acc = 0x8000 >> 1;
      for (int j = 0; j < 5; j++) acc = ox_add_sat(acc, ox_mult16x16(fp1[256 * j + k], c[k + 128 * j]));
      for (int j = 0; j < 5; j++) acc = ox_add_sat(acc, ox_mult16x16(fp2[128 + 256 * j + k], c[k + 64 + 128 * j]));
      out[ch_fac * k] = (i16)(ox_shl32_sat(acc, 1) >> 16);
    }
    off -= 128;
    if (off < 0) off += 1280;
    fpos += 64;
    if (fpos == 640) fpos = 0;
  }
  *drc_offset = off;
  *filter_pos = fpos;
}

/* decoder/generic/ixheaacd_qmf_dec_generic.c:590-741 (+ :468-526 fwd_modulation, :528-588 winadd) — complex 32-band
 * analysis of 1024 core samples into matrix[32][128] (re at +0..31, im at +64..95) */
int xo_anal_qmffilt_hq(const uint8_t *qrom, const i16 *time_in, int ch_fac, i16 *states, i32 *pos_io, i32 *fpos_io,
                       int usb, i32 *matrix) {
  const i16 *qmf_c = T16(XO_QROM_QMF_C);
  const i16 *tcos = T16(XO_QROM_TCOSSIN_L32);
  int pos = *pos_io;
  int f1 = *fpos_io, f2 = f1 + 64;
  for (int i = 0; i < 32; i++) {
    i32 buf[64];
    for (int k = 0; k < 32; k++) states[pos + 31 - k] = time_in[ch_fac * (32 * i + k)];
    const i16 *fp1 = states + ((i & 1) ? 32 : 0), *fp2 = states + ((i & 1) ? 0 : 32);
    for (int n = 0; n < 32; n++) {
      i32 a = ox_mult16x16(fp1[n], qmf_c[f1 + 2 * n]);
      i32 b = ox_mult16x16(fp2[n], qmf_c[f2 + 2 * n]);
      for (int j = 1; j < 5; j++) {
        a = ox_add_sat(a, ox_mult16x16(fp1[n + 64 * j], qmf_c[f1 + 2 * (n + 64 * j)]));
        b = ox_add_sat(b, ox_mult16x16(fp2[n + 64 * j], qmf_c[f2 + 2 * (n + 64 * j)]));
      }
      buf[n] = a;
      buf[n + 32] = b;
    }
    pos -= 32;
    if (pos < 0) pos = 288;
    { /* :696-718 — the two coefficient pointers leap-frog by 128 and wrap after 640 */
      int n1 = f2 + 64, n2 = f1 + 64;
      f1 = n1;
      f2 = n2;
      if (f2 > 640) {
        f1 = 0;
        f2 = 64;
      }
    }
    i32 *re = matrix + 128 * i, *im = re + 64;
    for (int k = 0; k < 32; k++) { /* :480-487 */
      i32 t1 = ox_shr32(buf[k], 4), t2 = ox_shr32(buf[63 - k], 4);
      re[k] = ox_sub_sat(t1, t2);
      im[k] = ox_add_sat(t1, t2);
    }
    xo_cos_sin_mod(qrom, re, 32);
    for (int k = 0; k < usb; k++) { /* :499-513 (lsb = 0 for the analysis bank) */
      i16 ch = tcos[2 * k], sh = tcos[2 * k + 1];
      i32 r = re[k], m = im[k];
      re[k] = ox_add_sat(ox_mul32x16_shl(r, ch), ox_mul32x16_shl(m, sh));
      im[k] = ox_sub_sat(ox_mul32x16_shl(m, ch), ox_mul32x16_shl(r, sh));
    }
  }
  *pos_io = pos;
  *fpos_io = f1;
  return -8; /* lb_scale for the HQ path, :635 */
}

void xo_synt_qmffilt_hq_batch(const uint8_t *qrom, i32 *matrix, i16 *fs, i32 *drc_offset, i32 *filter_pos,
                              const i32 *sf, const i32 *lsb, const i32 *usb, i16 *time_out, int n) {
  for (int u = 0; u < n; u++)
    xo_synt_qmffilt_hq(qrom, matrix + (size_t)u * 4096, fs + (size_t)u * 1280, drc_offset + u, filter_pos + u,
                       sf + 4 * u, lsb[u], usb[u], 6, time_out + (size_t)u * 2048, 1);
}

void xo_anal_qmffilt_hq_batch(const uint8_t *qrom, const i16 *time_in, i16 *states, i32 *pos, i32 *filter_pos,
                              const i32 *usb, i32 *matrix, int n) {
  for (int u = 0; u < n; u++)
    xo_anal_qmffilt_hq(qrom, time_in + (size_t)u * 1024, 1, states + (size_t)u * 320, pos + u, filter_pos + u, usb[u],
                       matrix + (size_t)u * 4096);
}

/* =====================================================================================================
 * Low-power (real-valued) SBR filterbanks: what the reference runs for stereo HE-AACv1 in its fixed-point path
 * (low_pow_flag = 1, decoder/ixheaacd_sbrdecoder.c:408-419).
 * ===================================================================================================== */

/* generic:63-239 — DCT-III of one slot: in[64] (window-add output, destroyed) -> out[0..31]; out needs 32 words */
static void dct3_32(const uint8_t *qrom, i32 *in, i32 *out) {
  const i16 *tw = T16(XO_QROM_DCT23_TW) + 4, *post = T16(XO_QROM_POST_FFT);
  int f = 49, r = 47, o = 0;
  i32 t0, t1, t2, t3, u0, u1, u2, u3, v4, v5;
  i16 re, im;
  out[o++] = in[48] >> 7;
  out[o++] = 0;
  for (int n = 1; n < 16; n++) {
    t0 = in[f++];
    t1 = in[r--];
    t0 = ox_add_sat(ox_shr32(t0, 7), ox_shr32(t1, 7));
    t2 = in[f - 33];
    t3 = in[r - 31];
    t1 = ox_sub_sat(ox_shr32(t2, 7), ox_shr32(t3, 7));
    re = tw[0];
    im = tw[1];
    tw += 4;
    out[o++] = ox_add(ox_mul32x16(t0, re), ox_mul32x16(t1, im));
    out[o++] = ox_add(ox_sub(0, ox_mul32x16(t1, re)), ox_mul32x16(t0, im));
  }
  re = tw[0];
  im = tw[1];
  t1 = in[r--];
  t0 = in[r - 31];
  t1 = ox_sub_sat(ox_shr32(t1, 7), ox_shr32(t0, 7));
  t0 = t1;
  u2 = ox_add(ox_mul32x16(t0, re), ox_mul32x16(t1, im));
  u3 = ox_add(ox_sub(0, ox_mul32x16(t1, re)), ox_mul32x16(t0, im));
  int pf = 0, pr = 31;
  u0 = out[0];
  u1 = out[1];
  t0 = ox_sub(ox_sub(0, u1), u3);
  t1 = ox_sub(u0, u2);
  u0 = ox_add(ox_add(u0, u2), t0);
  u1 = ox_add(ox_sub(u1, u3), t1);
  out[pf++] = u0 >> 1;
  out[pf++] = u1 >> 1;
  const i16 *tf = post + 2, *tr = post + 14;
  for (int n = 1; n <= 8; n++) {
    const int last = n == 8;
    u0 = out[pf];
    u1 = out[pf + 1];
    u3 = out[pr];
    u2 = out[pr - 1];
    re = last ? (i16)(-*tr) : *tr;
    tr -= 2;
    im = *tf;
    tf += 2;
    t0 = ox_sub(u0, u2);
    t1 = ox_add(u0, u2);
    t2 = ox_add(u1, u3);
    t3 = ox_sub(u1, u3);
    if (!last) {
      v4 = ox_add(ox_mul32x16(t0, re), ox_mul32x16(t2, im));
      v5 = ox_add(ox_sub(0, ox_mul32x16(t2, re)), ox_mul32x16(t0, im));
    } else {
      v4 = ox_sub(ox_mul32x16(t0, re), ox_mul32x16(t2, im));
      v5 = ox_add(ox_mul32x16(t2, re), ox_mul32x16(t0, im));
    }
    t1 >>= 1;
    t3 >>= 1;
    if (!last) {
      out[pf++] = ox_sub(t1, v4);
      out[pf++] = ox_add(t3, v5);
      out[pr--] = ox_add(ox_sub(0, t3), v5);
      out[pr--] = ox_add(t1, v4);
    } else {
      out[pf++] = ox_add(t1, v4);
      out[pf++] = ox_add(t3, v5);
    }
  }
  radix4_stage(T16(XO_QROM_W16), out, 1, 4);
  post_radix4_16(in, out, T32(XO_QROM_DIGREV4_16));
  out[0] = in[0];
  out[2] = in[1];
  int po = 2, p1 = 18;
  pf = 1;
  pr = 30;
  for (int k = 7; k != 0; k--) {
    i32 a = in[po++], b = in[po++];
    out[pf] = b; pf += 2;
    out[pf] = a; pf += 2;
    a = in[p1++];
    b = in[p1++];
    out[pr] = b; pr -= 2;
    out[pr] = a; pr -= 2;
  }
  {
    i32 a = in[po++], b = in[po++];
    out[pf] = b; pf += 2;
    out[pf] = a;
  }
}

void xo_dct3_32(const uint8_t *qrom, i32 *in, i32 *out) { dct3_32(qrom, in, out); }

/* decoder/generic/ixheaacd_qmf_dec_generic.c:590-741 with low_pow_flag = 1: real 32-band analysis of 1024 core samples
 * into matrix[32][64] (bands 0..31 of each 64-word row written).  Returns lb_scale (-10, :631). */
int xo_anal_qmffilt_lp(const uint8_t *qrom, const i16 *time_in, int ch_fac, i16 *states, i32 *pos_io, i32 *fpos_io,
                       i32 *matrix) {
  const i16 *qmf_c = T16(XO_QROM_QMF_C);
  int pos = *pos_io;
  int f1 = *fpos_io, f2 = f1 + 64;
  for (int i = 0; i < 32; i++) {
    i32 buf[128];
    for (int k = 0; k < 32; k++) states[pos + 31 - k] = time_in[ch_fac * (32 * i + k)];
    const i16 *fp1 = states + ((i & 1) ? 32 : 0), *fp2 = states + ((i & 1) ? 0 : 32);
    for (int n = 0; n < 32; n++) {
      i32 a = ox_mult16x16(fp1[n], qmf_c[f1 + 2 * n]);
      i32 b = ox_mult16x16(fp2[n], qmf_c[f2 + 2 * n]);
      for (int j = 1; j < 5; j++) {
        a = ox_add_sat(a, ox_mult16x16(fp1[n + 64 * j], qmf_c[f1 + 2 * (n + 64 * j)]));
        b = ox_add_sat(b, ox_mult16x16(fp2[n + 64 * j], qmf_c[f2 + 2 * (n + 64 * j)]));
      }
      buf[n] = a;
      buf[n + 32] = b;
    }
    pos -= 32;
    if (pos < 0) pos = 288;
    {
      int n1 = f2 + 64, n2 = f1 + 64;
      f1 = n1;
      f2 = n2;
      if (f2 > 640) { f1 = 0; f2 = 64; }
    }
    dct3_32(qrom, buf, matrix + 64 * i);
  }
  *pos_io = pos;
  *fpos_io = f1;
  return -10;
}

/* decoder/ixheaacd_qmf_dec.c:72-211 + generic:241-257 — DCT-II of one slot straight into 128 WORD16 state samples.
 * x[64] is destroyed; fs points at the slot's state block (generic:851-867: dct2_64(.., filter_states + 32), [96] = 0) */
static void inv_modulation_lp(const uint8_t *qrom, i32 *x, i16 *fs) {
  i32 X[64];
  for (int n = 0; n < 32; n++) { X[n] = x[2 * n]; X[63 - n] = x[2 * n + 1]; } /* pretwdct2 */
  radix4_stage(T16(XO_QROM_W32), X, 1, 8);
  radix4_stage(T16(XO_QROM_W32) + 48, X, 4, 2);
  post_radix2_32(x, X, T32(XO_QROM_DIGREV2_32));
  { /* fftposttw, qmf_dec.c:107-159 */
    const i16 *tf = T16(XO_QROM_POST_FFT) + 1, *tr = T16(XO_QROM_POST_FFT) + 15;
    int pf = 0, pr = 63;
    x[0] = ox_shl1(x[0]);
    x[1] = ox_shl1(x[1]);
    pf = 2;
    for (int k = 1; k <= 16; k++) {
      i32 t0 = x[pf], t1 = x[pf + 1], t3 = x[pr], t2 = x[pr - 1];
      i32 in2 = ox_sub_sat(t3, t1), in1 = ox_add_sat(t3, t1);
      t1 = ox_sub_sat(t0, t2);
      t3 = ox_add_sat(t0, t2);
      i16 re = *tf++, im = *tr--;
      i32 v1 = ox_shl1(ox_sub(ox_mul32x16(in1, re), ox_mul32x16(t1, im)));
      i32 v2 = ox_shl1(ox_add(ox_mul32x16(t1, re), ox_mul32x16(in1, im)));
      x[pf++] = ox_add_sat(t3, v1);
      x[pf++] = ox_add_sat(in2, v2);
      x[pr--] = ox_sub_sat(v2, in2);
      x[pr--] = ox_sub_sat(t3, v1);
    }
  }
  { /* posttwdct2, qmf_dec.c:161-211: out_fwd = fs + 32 */
    i16 *of = fs + 32, *orv = fs + 32 + 63, *or2 = fs + 31, *of2 = fs + 32 + 65;
    const i16 *tw = T16(XO_QROM_DCT23_TW) + 2;
    int p = 0;
    i32 ore = x[p++], oim = x[p++];
    i64 s = ((i64)ore + (i64)oim) >> 1;
    i32 ore1 = s >= OX_MAX32 ? OX_MAX32 : (s <= OX_MIN32 ? OX_MIN32 : (i32)s);
    *of++ = ox_round16(ox_shl32(ore1, 4));
    i32 last = ox_sub_sat(ore, oim);
    for (int k = 30; k >= 0; k--) {
      i32 ire = x[p++], iim = x[p++];
      i16 re = *tw++, im = *tw++;
      ore = ox_sub_sat(ox_mul32x16(ire, re), ox_mul32x16(iim, im));
      oim = ox_add_sat(ox_mul32x16(iim, re), ox_mul32x16(ire, im));
      i16 r1 = ox_round16(ox_shl32(ore, 4)), i1 = ox_round16(ox_shl32(oim, 4)), i2 = ox_neg16(i1);
      *of++ = r1;
      *or2-- = r1;
      *orv-- = i1;
      *of2++ = i2;
    }
    i16 r1 = ox_round16(ox_shl32(ox_mul32x16(last, *tw), 4));
    *of++ = r1;
    *or2-- = r1;
  }
  fs[96] = 0;
}

/* decoder/ixheaacd_qmf_dec.c:811-1129 with low_pow_flag = 1: real 64-band synthesis.  matrix [32][64] in place. */
void xo_synt_qmffilt_lp(const uint8_t *qrom, i32 *matrix, i16 *fs, i32 *drc_offset, i32 *filter_pos, const i32 *sf,
                        int lsb, int usb, int split, i16 *time_out, int ch_fac) {
  const i16 *qmf_c = T16(XO_QROM_QMF_C);
  const int shifts[3] = {(sf[3] - sf[0]) - 4, (sf[3] - sf[1]) - 4, (sf[3] - sf[2]) - 4}; /* ov_lb, lb, hb: :905-907 */
  for (int l = 0; l < 32; l++)
    for (int k = 0; k < usb; k++) {
      int sh = k < lsb ? ((shifts[0] == shifts[1] || l >= split) ? shifts[1] : shifts[0]) : shifts[2];
      if (k < lsb && shifts[0] == shifts[1]) sh = shifts[0];
      if (sh > 31) sh = 31;
      if (sh < -31) sh = -31;
      i32 *p = matrix + 64 * l + k;
      if (sh > 0) *p = ox_lsl(*p, sh);
      else if (sh < 0) *p = *p >> -sh;
    }
  int off = *drc_offset, fpos = *filter_pos;
  for (int i = 0; i < 32; i++) {
    inv_modulation_lp(qrom, matrix + 64 * i, fs + off);
    const i16 *fp1 = fs + ((i & 1) ? 64 : 0), *fp2 = fs + ((i & 1) ? 0 : 64);
    const i16 *c = qmf_c + fpos;
    i16 *out = time_out + ch_fac * 64 * i;
    for (int k = 0; k < 64; k++) { /* generic:1508-1542 with shift = 2 */
      i32 acc = 0x8000 >> 2;
      for (int j = 0; j < 5; j++) acc = ox_add_sat(acc, ox_mult16x16(fp1[256 * j + k], c[k + 128 * j]));
      for (int j = 0; j < 5; j++) acc = ox_add_sat(acc, ox_mult16x16(fp2[128 + 256 * j + k], c[k + 64 + 128 * j]));
      out[ch_fac * k] = (i16)(ox_shl32_sat(acc, 2) >> 16);
    }
    off -= 128;
    if (off < 0) off += 1280;
    fpos += 64;
    if (fpos == 640) fpos = 0;
  }
  *drc_offset = off;
  *filter_pos = fpos;
}
