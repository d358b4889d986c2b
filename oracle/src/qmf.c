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
