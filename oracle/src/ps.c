/*
 * oracle/src/ps.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain-C restatement of libxaac's fixed-point parametric-stereo decoder as the synthesis stage
 * runs it: block-floating-point rescale of the PS state (ixheaacd_init_ps_scale), and per QMF slot the hybrid
 * analysis of the three lowest bands, the transient-steered all-pass / delay decorrelator and the 2x2 rotation,
 * followed by the two 64-band syntheses (left with the PS scale, right from the decorrelated matrix).
 * Cites reference lines (paths relative to /root/reference).  Pinned through the whole-stage records of
 * tests/golden/sbrdec_tapped.npz (HE-AACv2 decodes of the compiled reference).
 */
#include <string.h>
#include "fixmath.h"
#include "xaac_oracle.h"

#define T16P(off) ((const i16 *)(ps_rom + 2 * (off)))

/* decoder/ixheaacd_ps_dec.c:188-210 with ixheaacd_get_ps_scale (:125-186) and ixheaacd_scale_ps_states
 * (decoder/ixheaacd_thumb_ps_dec.c:101-181) */
static i32 or_pairs(const i16 *p, int pairs) {
  i32 mx = 0;
  for (int i = 0; i < 2 * pairs; i++) mx |= ox_abs_nrm(p[i]);
  return mx;
}
static void shl16v(i16 *p, int n, int s) { for (int i = 0; i < n; i++) p[i] = ox_sat16((i32)p[i] << s); }
static void shr16v(i16 *p, int n, int s) { for (int i = 0; i < n; i++) p[i] = (i16)(p[i] >> s); }
static void shl32v(i32 *p, int n, int s) { for (int i = 0; i < n; i++) p[i] = ox_shl32_sat(p[i], s); }
static void shr32v(i32 *p, int n, int s) { for (int i = 0; i < n; i++) p[i] = ox_shr32(p[i], s); }

static void ps_init_scale(i16 *ps, i16 *sf, const i16 *delay_ser) {
  i16 *ap = ps + XO_PS_ST_AP, *ld = ps + XO_PS_ST_LD, *sd = ps + XO_PS_ST_SD, *ser = ps + XO_PS_ST_SER;
  i16 *sub = ps + XO_PS_ST_SUB, *subser = ps + XO_PS_ST_SUB_SER, *idx = ps + XO_PS_ST_IDX;
  i32 *peak = (i32 *)(ps + XO_PS_ST_PEAK), *hyb = (i32 *)(ps + XO_PS_ST_HYB);
  i32 mx = 0;
  for (int m = 0; m < 2; m++) mx |= or_pairs(ap + 64 * m + 6, 20);
  mx |= or_pairs(ld, 14 * 12);
  mx |= or_pairs(sd, 29);
  mx |= or_pairs(sub, 32);
  for (int i = 0; i < 3; i++)
    for (int m = 0; m < delay_ser[i]; m++) mx |= or_pairs(ser + 192 * m + 64 * i + 6, 20);
  mx |= or_pairs(subser, 240);
  mx = (i32)((u32)mx << 16);
  for (int i = 0; i < 72; i++) mx |= ox_abs_nrm(hyb[i]);
  int reserve = ox_pnorm32(mx);

  idx[XO_PS_IDX_SCALE] = (i16)(idx[XO_PS_IDX_SCALE] + reserve);
  i16 t = sf[XO_SF_LB] < sf[XO_SF_OV_LB] ? sf[XO_SF_LB] : sf[XO_SF_OV_LB];
  if (sf[XO_SF_HB] < t) t = sf[XO_SF_HB];
  if (idx[XO_PS_IDX_SCALE] < t) t = idx[XO_PS_IDX_SCALE];
  sf[XO_SF_PS] = (i16)(t - 1);
  int change = (sf[XO_SF_PS] - idx[XO_PS_IDX_SCALE]) + reserve;
  i16 scale = (i16)change;
  if (scale > 0) {
    int s1 = scale > 15 ? 15 : scale;
    for (int m = 0; m < 2; m++) shl16v(ap + 64 * m + 6, 40, s1);
    shl16v(ld, 336, s1);
    shl16v(sd, 58, s1);
    shl16v(sub, 64, s1);
    shl16v(subser, 480, s1);
    for (int i = 0; i < 3; i++)
      for (int m = 0; m < delay_ser[i]; m++) shl16v(ser + 192 * m + 64 * i + 6, 40, s1);
    shl32v(hyb, 72, scale);
    shl32v(peak, 60, (i16)(scale + scale));
  } else if (scale != 0) {
    scale = (i16)-scale;
    for (int m = 0; m < 2; m++) shr16v(ap + 64 * m + 6, 40, scale);
    shr16v(ld, 336, scale);
    shr16v(sd, 58, scale);
    shr16v(sub, 64, scale);
    shr16v(subser, 480, scale);
    for (int i = 0; i < 3; i++)
      for (int m = 0; m < delay_ser[i]; m++) shr16v(ser + 192 * m + 64 * i + 6, 40, scale);
    shr32v(hyb, 72, scale);
    shr32v(peak, 60, (i16)(scale + scale));
  }
  idx[XO_PS_IDX_SCALE] = sf[XO_SF_PS];
}

/* decoder/ixheaacd_dsp_fft32x32s.c:34-117 */
static void fft8(const i32 *y, i32 *real, i32 *imag) {
  i32 a0, a1, a2, a3, a00, a10, a20, a30, vr, vi, x[16];
  a00 = ox_add_sat(y[0], y[8]); a0 = ox_sub_sat(y[0], y[8]);
  a20 = ox_add_sat(y[1], y[9]); a3 = ox_sub_sat(y[1], y[9]);
  a10 = ox_add_sat(y[4], y[12]); a2 = ox_sub_sat(y[4], y[12]);
  a30 = ox_add_sat(y[5], y[13]); a1 = ox_sub_sat(y[5], y[13]);
  x[0] = ox_add_sat(a00, a10); x[4] = ox_sub_sat(a00, a10);
  x[1] = ox_add_sat(a20, a30); x[5] = ox_sub_sat(a20, a30);
  x[2] = ox_sub_sat(a0, a1); x[6] = ox_add_sat(a0, a1);
  x[3] = ox_add_sat(a3, a2); x[7] = ox_sub_sat(a3, a2);
  a00 = ox_add_sat(y[2], y[10]); a0 = ox_sub_sat(y[2], y[10]);
  a20 = ox_add_sat(y[3], y[11]); a3 = ox_sub_sat(y[3], y[11]);
  a10 = ox_add_sat(y[6], y[14]); a2 = ox_sub_sat(y[6], y[14]);
  a30 = ox_add_sat(y[7], y[15]); a1 = ox_sub_sat(y[7], y[15]);
  x[8] = ox_add_sat(a00, a10); x[12] = ox_sub_sat(a00, a10);
  x[9] = ox_add_sat(a20, a30); x[13] = ox_sub_sat(a20, a30);
  x[10] = ox_sub_sat(a0, a1); x[14] = ox_add_sat(a0, a1);
  x[11] = ox_add_sat(a3, a2); x[15] = ox_sub_sat(a3, a2);
  real[0] = ox_add_sat(x[0], x[8]);
  imag[0] = ox_add_sat(x[1], x[9]);
  a00 = ox_sub_sat(x[0], x[8]);
  a10 = ox_sub_sat(x[1], x[9]);
  a0 = ox_sub_sat(x[4], x[13]);
  a1 = ox_add_sat(x[5], x[12]);
  real[4] = ox_add_sat(x[4], x[13]);
  imag[4] = ox_sub_sat(x[5], x[12]);
#define MSS(a) ox_mul32x16_shl((a), 0x5A82) /* mult32x16in32_shl_sat only saturates for a multiplier of 0x8000 */
  vr = MSS(ox_sub_sat(x[10], x[11]));
  vi = MSS(ox_add_sat(x[10], x[11]));
  real[1] = ox_add_sat(x[2], vr);
  imag[1] = ox_add_sat(x[3], vi);
  a2 = ox_sub_sat(x[2], vr);
  a3 = ox_sub_sat(x[3], vi);
  real[2] = ox_add_sat(a0, a2);
  imag[2] = ox_add_sat(a1, a3);
  vr = MSS(ox_add_sat(x[14], x[15]));
  vi = MSS(ox_sub_sat(x[14], x[15]));
#undef MSS
  a20 = ox_sub_sat(x[6], vr);
  a30 = ox_add_sat(x[7], vi);
  real[3] = ox_add_sat(a00, a20);
  imag[3] = ox_add_sat(a10, a30);
  real[5] = ox_add_sat(x[6], vr);
  imag[5] = ox_sub_sat(x[7], vi);
}

#define M(a, c) ox_mul32x16((a), (c))
#define S1(a) ox_shl32((a), 1)

/* decoder/ixheaacd_hybrid.c:96-212.  re/im: 13 delayed samples each. */
static void filt_8_ch(const i32 *re, const i32 *im, i32 *hr, i32 *hi, const i16 *p) {
  const i16 tcos = 0x7642, tsin = 0x30fc, tcom = 0x5a82;
  i32 real, imag, cum[16];
  real = S1(ox_add_sat(M(re[0], p[0]), M(re[8], p[8])));
  imag = S1(ox_add_sat(M(im[0], p[0]), M(im[8], p[8])));
  cum[12] = S1(M(ox_add_sat(imag, real), tcom));
  cum[13] = S1(M(ox_sub_sat(imag, real), tcom));
  real = S1(ox_add_sat(M(re[1], p[1]), M(re[9], p[9])));
  imag = S1(ox_add_sat(M(im[1], p[1]), M(im[9], p[9])));
  cum[10] = S1(ox_add_sat(M(imag, tcos), M(real, tsin)));
  cum[11] = S1(ox_sub_sat(M(imag, tsin), M(real, tcos)));
  cum[9] = S1(M(ox_sub_sat(re[2], re[10]), p[10]));
  cum[8] = S1(M(ox_sub_sat(im[2], im[10]), p[2]));
  real = S1(ox_add_sat(M(re[3], p[3]), M(re[11], p[11])));
  imag = S1(ox_add_sat(M(im[3], p[3]), M(im[11], p[11])));
  cum[6] = S1(ox_sub_sat(M(imag, tcos), M(real, tsin)));
  cum[7] = S1(ox_neg_sat(ox_add_sat(M(imag, tsin), M(real, tcos))));
  real = S1(ox_add_sat(M(re[4], p[4]), M(re[12], p[12])));
  imag = S1(ox_add_sat(M(im[4], p[4]), M(im[12], p[12])));
  cum[4] = S1(M(ox_sub_sat(imag, real), tcom));
  cum[5] = S1(M(ox_neg_sat(ox_add_sat(imag, real)), tcom));
  real = S1(M(re[5], p[5]));
  imag = S1(M(im[5], p[5]));
  cum[2] = S1(ox_sub_sat(M(real, tcos), M(imag, tsin)));
  cum[3] = S1(ox_add_sat(M(real, tsin), M(imag, tcos)));
  cum[0] = S1(M(re[6], p[6]));
  cum[1] = S1(M(im[6], p[6]));
  real = S1(M(re[7], p[7]));
  imag = S1(M(im[7], p[7]));
  cum[14] = S1(ox_add_sat(M(imag, tsin), M(real, tcos)));
  cum[15] = S1(ox_sub_sat(M(imag, tcos), M(real, tsin)));
  fft8(cum, hr, hi);
}

/* decoder/ixheaacd_hybrid.c:51-94, one component (called for the real and the imaginary delay line) */
static void filt_2_ch(const i32 *q, i32 *h, const i16 *p2_6) {
  i32 cum0 = q[6] >> 1, cum1 = 0;
  for (int j = 0; j < 6; j++) cum1 = ox_add_sat(cum1, M(q[1 + 2 * j], p2_6[j]));
  cum1 = S1(cum1);
  h[0] = ox_add_sat(cum0, cum1);
  h[1] = ox_sub_sat(cum0, cum1);
}

/* decoder/ixheaacd_hybrid.c:214-285.  row6: QMF row of slot + 6 (re[64] | im[64]); hyb: left_re[16] | left_im[16] */
static void hybrid_analysis(const i32 *row6, i32 *hyb, i32 *qbuf, int scale, const uint8_t *ps_rom) {
  const i16 *resol = T16P(XO_PSROM_HYB_RESOL);
  int off = 0;
  for (int band = 0; band < 3; band++) {
    i32 wre[13], wim[13];
    i32 *bre = qbuf + 24 * band, *bim = bre + 12;
    memcpy(wre, bre, 12 * sizeof(i32));
    memcpy(wim, bim, 12 * sizeof(i32));
    memmove(bre, bre + 1, 11 * sizeof(i32));
    memmove(bim, bim + 1, 11 * sizeof(i32));
    i32 tr = row6[band], ti = row6[band + 64];
    if (scale < 0) { tr = ox_shl32(tr, -scale); ti = ox_shl32(ti, -scale); }
    else { tr = ox_shr32(tr, scale); ti = ox_shr32(ti, scale); }
    wre[12] = bre[11] = tr;
    wim[12] = bim[11] = ti;
    if (resol[band] == 2) {
      filt_2_ch(wre, hyb + off, T16P(XO_PSROM_P2_6));
      filt_2_ch(wim, hyb + 16 + off, T16P(XO_PSROM_P2_6));
      off += 2;
    } else if (resol[band] == 8) {
      filt_8_ch(wre, wim, hyb + off, hyb + 16 + off, T16P(XO_PSROM_P8_13));
      off += 6;
    }
  }
}

/* decoder/ixheaacd_ps_dec.c:212-234 */
static i32 divide16_pos(i32 op1, i32 op2) {
  int nrm = ox_norm32(op2);
  u32 u = (u32)op1 << nrm, v = (u32)op2 << nrm;
  u &= 0xffff0000u;
  v &= 0xffff0000u;
  if (u != 0)
    for (int k = 16; k > 0; k--) {
      if (u >= v) u = ((u - v) << 1) + 1;
      else u <<= 1;
    }
  return (i32)u;
}

static inline i32 pw(i32 v) { return ox_mul32x16(v, (i16)(v >> 16)); }
static inline i16 rot_re(i16 r, i16 i, const i16 *f) {
  return (i16)(ox_sub_sat(ox_mult16x16(r, f[0]), ox_mult16x16(i, f[1])) >> 15);
}
static inline i16 rot_im(i16 r, i16 i, const i16 *f) {
  return (i16)(ox_add_sat(ox_mult16x16(r, f[1]), ox_mult16x16(i, f[0])) >> 15);
}

/* three serial all-pass links on one (sub)band (decoder/ixheaacd_ps_dec.c:274-312, 398-438).
 * d: delay-line base for link m at [192 m' + 64 m]-style addressing supplied by the caller through `at` */
typedef struct { i16 *p[3]; const i16 *fac[3]; i16 decay[3]; } ap_t;
static void allpass3(i16 *rin, i16 *iin, const ap_t *a) {
  i16 real_in = *rin, imag_in = *iin;
  for (int m = 0; m < 3; m++) {
    i16 r0 = a->p[m][0], i0 = a->p[m][1];
    i16 rt = rot_re(r0, i0, a->fac[m]), it = rot_im(r0, i0, a->fac[m]);
    rt = ox_sub16(rt, ox_mult16_shl(real_in, a->decay[m]));
    it = ox_sub16(it, ox_mult16_shl(imag_in, a->decay[m]));
    a->p[m][0] = ox_add16(real_in, ox_mult16_shl(rt, a->decay[m]));
    a->p[m][1] = ox_add16(imag_in, ox_mult16_shl(it, a->decay[m]));
    real_in = rt;
    imag_in = it;
  }
  *rin = real_in;
  *iin = imag_in;
}

/* decoder/ixheaacd_ps_dec.c:450-675 (with decorr_filter1 :236-337 and decorr_filter2 :339-448).
 * left: QMF row (re|im) of this slot, right: output row (re|im); hyb: left_re|left_im|right_re|right_im [16 each] */
static void decorrelation(i16 *ps, const i32 *left, i32 *right, i32 *hyb, const uint8_t *ps_rom) {
  i16 *idx = ps + XO_PS_ST_IDX;
  i32 *peak = (i32 *)(ps + XO_PS_ST_PEAK), *nrg_prev = peak + 20, *peak_prev = peak + 40;
  const i16 *borders = T16P(XO_PSROM_BORDERS_GROUP), *gshift = T16P(XO_PSROM_GROUP_SHIFT);
  const i32 *lre = hyb, *lim = hyb + 16;
  i32 *rre = hyb + 32, *rim = hyb + 48;
  const int usb = idx[XO_PS_IDX_USB];
  i32 power[20];
  i16 tr[21];

  power[0] = ox_add_sat(ox_add_sat(ox_add_sat(pw(lre[0]), pw(lim[0])), pw(lre[5])), pw(lim[5]));
  power[1] = ox_add_sat(ox_add_sat(ox_add_sat(pw(lre[4]), pw(lim[4])), pw(lre[1])), pw(lim[1]));
  for (int gr = 4, bin = 2; gr < 10; gr++, bin++) {
    int sb = borders[gr];
    power[bin] = ox_add_sat(pw(lre[sb]), pw(lim[sb]));
  }
  for (int sband = 3, bin = 8; sband < 9; sband++, bin++) power[bin] = ox_add_sat(pw(left[sband]), pw(left[64 + sband]));
  for (int gr = 16, bin = 14; gr < 22; gr++, bin++) {
    i32 accu = 0;
    int mxs = usb < borders[gr + 1] ? usb : borders[gr + 1];
    for (int sband = borders[gr]; sband < mxs; sband++) {
      i32 t = ox_add_sat(pw(left[sband]), pw(left[64 + sband]));
      accu = ox_add_sat(accu, t >> gshift[gr - 16]);
    }
    power[bin] = accu;
  }
  for (int bin = 0; bin < 20; bin++) {
    i32 p = ox_shl32(power[bin], 1);
    if (p < 0) p = 0;
    peak[bin] = ox_mul32x16_shl(peak[bin], 0x620a);
    if (p > peak[bin]) peak[bin] = p;
    i32 pd = ox_add_sat(ox_mul32x16_shl(peak_prev[bin], 0x6000), ox_sub_sat(peak[bin], p) >> 2);
    peak_prev[bin] = pd;
    i32 nrg = ox_add_sat(ox_mul32x16_shl(nrg_prev[bin], 0x6000), p >> 2);
    nrg_prev[bin] = nrg;
    pd = ox_add_sat(pd, pd >> 1);
    tr[bin] = pd <= nrg ? 0x7fff : (i16)divide16_pos(nrg, pd);
  }

  const int di0 = idx[XO_PS_IDX_DELAY];
  const i16 *decay_ser = T16P(XO_PSROM_REV_DECAY);
  { /* filter 1: the 10 hybrid sub-subbands */
    i16 *dsub = ps + XO_PS_ST_SUB + 32 * di0;
    const i16 *fac = T16P(XO_PSROM_FRAC_SUB), *facs = T16P(XO_PSROM_FRAC_SUB_SER);
    const i16 *h2b = T16P(XO_PSROM_HYB_TO_BIN);
    for (int sb = 0; sb < 10; sb++) {
      i16 r0 = dsub[2 * sb], i0 = dsub[2 * sb + 1];
      i16 rin = rot_re(r0, i0, fac + 2 * sb), iin = rot_im(r0, i0, fac + 2 * sb);
      dsub[2 * sb] = ox_round16(lre[sb]);
      dsub[2 * sb + 1] = ox_round16(lim[sb]);
      ap_t a;
      for (int m = 0; m < 3; m++) {
        a.p[m] = ps + XO_PS_ST_SUB_SER + 96 * idx[XO_PS_IDX_SER + m] + 32 * m + 2 * sb;
        a.fac[m] = facs + 32 * m + 2 * sb;
        a.decay[m] = decay_ser[m];
      }
      allpass3(&rin, &iin, &a);
      rre[sb] = ox_shl32(ox_mult16x16(rin, tr[h2b[sb]]), 1);
      rim[sb] = ox_shl32(ox_mult16x16(iin, tr[h2b[sb]]), 1);
    }
  }
  tr[20] = 0;
  { /* filter 2: QMF bands 3..22 */
    i16 *dap = ps + XO_PS_ST_AP + 64 * di0;
    const i16 *fac = T16P(XO_PSROM_FRAC_QMF), *facs = T16P(XO_PSROM_FRAC_QMF_SER);
    const i16 *d2b = T16P(XO_PSROM_DELAY_TO_BIN), *dsf = T16P(XO_PSROM_DECAY_SF);
    for (int sb = 3, di = 9; sb < 23; sb++, di += 3) {
      i16 r0 = dap[2 * sb], i0 = dap[2 * sb + 1];
      i16 rin = rot_re(r0, i0, fac + 2 * sb), iin = rot_im(r0, i0, fac + 2 * sb);
      dap[2 * sb] = ox_round16(left[sb]);
      dap[2 * sb + 1] = ox_round16(left[64 + sb]);
      ap_t a;
      for (int m = 0; m < 3; m++) {
        a.p[m] = ps + XO_PS_ST_SER + 192 * idx[XO_PS_IDX_SER + m] + 64 * m + 2 * sb;
        a.fac[m] = facs + 64 * m + 2 * sb;
        a.decay[m] = dsf[di + m];
      }
      allpass3(&rin, &iin, &a);
      right[sb] = ox_shl32(ox_mult16x16(rin, tr[d2b[sb]]), 1);
      right[64 + sb] = ox_shl32(ox_mult16x16(iin, tr[d2b[sb]]), 1);
    }
  }
  { /* :596-645 — plain delays: 14 slots for bands 23..34, 1 slot above */
    int mxs = (i16)usb < borders[21] ? (i16)usb : borders[21];
    i16 *d = ps + XO_PS_ST_LD + 24 * idx[XO_PS_IDX_DELAY_LONG];
    for (int sband = borders[20]; sband < mxs; sband++, d += 2) {
      i16 r = d[0], i = d[1];
      d[0] = ox_round16(left[sband]);
      d[1] = ox_round16(left[64 + sband]);
      right[sband] = ox_shl32(ox_mult16x16(r, tr[18]), 1);
      right[64 + sband] = ox_shl32(ox_mult16x16(i, tr[18]), 1);
    }
    idx[XO_PS_IDX_DELAY_LONG] = ox_add16(idx[XO_PS_IDX_DELAY_LONG], 1);
    if (idx[XO_PS_IDX_DELAY_LONG] >= 14) idx[XO_PS_IDX_DELAY_LONG] = 0;
    d = ps + XO_PS_ST_SD;
    mxs = (i16)usb < borders[22] ? (i16)usb : borders[22];
    for (int sband = borders[21]; sband < mxs; sband++, d += 2) {
      i16 r = d[0], i = d[1];
      d[0] = ox_round16(left[sband]);
      d[1] = ox_round16(left[64 + sband]);
      right[sband] = ox_shl32(ox_mult16x16(r, tr[19]), 1);
      right[64 + sband] = ox_shl32(ox_mult16x16(i, tr[19]), 1);
    }
  }
  for (int sband = usb; sband < 64; sband++) right[sband] = right[64 + sband] = 0;
  idx[XO_PS_IDX_DELAY] = (i16)(idx[XO_PS_IDX_DELAY] + 1);
  if (idx[XO_PS_IDX_DELAY] >= 2) idx[XO_PS_IDX_DELAY] = 0;
  const i16 *dser = T16P(XO_PSROM_REV_DELAY);
  for (int m = 0; m < 3; m++) {
    idx[XO_PS_IDX_SER + m] = (i16)(idx[XO_PS_IDX_SER + m] + 1);
    if (idx[XO_PS_IDX_SER + m] >= dser[m]) idx[XO_PS_IDX_SER + m] = 0;
  }
}

/* decoder/ixheaacd_ps_dec.c:677-712 */
static i16 cos512(i32 phi, const i16 *tab) {
  i32 a = phi == OX_MIN32 ? OX_MAX32 : (phi < 0 ? -phi : phi);
  int index = ox_round16(a) & 0x3ff;
  return index < 512 ? tab[512 - index] : (i16)(-tab[index - 512]);
}
static i16 sin512(i32 phi, const i16 *tab) {
  int index = ox_round16(phi);
  if (index < 0) {
    index = (-index) & 0x3ff;
    return index < 512 ? (i16)(-tab[index]) : (i16)(-tab[1024 - index]);
  }
  index &= 0x3ff;
  return index < 512 ? tab[index] : tab[1024 - index];
}

/* decoder/ixheaacd_ps_dec.c:714-854 */
static void init_rot_env(i16 *ps, const i16 *prm, int env, int usb, const uint8_t *ps_rom, const i16 *inv_int,
                         const i16 *trig) {
  i16 *idx = ps + XO_PS_ST_IDX, *hv = ps + XO_PS_ST_HVEC;
  i16 *h11v = hv, *h21v = hv + 48, *H11 = hv + 96, *H21 = hv + 144, *d11 = hv + 192, *d21 = hv + 240;
  const i32 rescale = (i32)((u32)0x0517cc1b << 1);
  if (env == 0) {
    int usb_prev = idx[XO_PS_IDX_USB];
    idx[XO_PS_IDX_USB] = (i16)usb;
    if (usb > usb_prev && usb_prev) {
      const i16 *dser = T16P(XO_PSROM_REV_DELAY);
      int o = usb < 20 ? usb : 20;
      if (o > usb_prev)
        for (int i = 0; i < 3; i++)
          for (int j = 0; j < dser[i]; j++)
            memset(ps + XO_PS_ST_SER + 192 * j + 64 * i + 2 * usb_prev, 0, sizeof(i16) * 2 * (o - usb_prev));
      int o1 = usb < 32 ? usb : 32;
      if (o1 >= o && o1 <= 12)
        for (int i = 0; i < 14; i++) memset(ps + XO_PS_ST_LD + 24 * i + 2 * o, 0, sizeof(i16) * 2 * (o1 - o));
      if (usb >= o1 && usb <= 16) memset(ps + XO_PS_ST_SD + 2 * o1, 0, sizeof(i16) * 2 * (usb - o1));
    }
  }
  const int fine = prm[XO_PS_PRM_IID_QUANT];
  const int steps = fine ? 15 : 7;
  const i16 *sfac = T16P(fine ? XO_PSROM_SCALE_FINE : XO_PSROM_SCALE);
  const i16 *alpha_tab = T16P(XO_PSROM_ALPHA), *g2b = T16P(XO_PSROM_GROUP_TO_BIN);
  const i16 *bp = prm + XO_PS_PRM_BORDER;
  i16 dl = ox_sub16_sat(bp[env + 1], bp[env]);
  i16 inv_len = inv_int[dl < 0 ? (i16)-dl : dl];
  const i16 *iid = prm + XO_PS_PRM_IID + 34 * env, *icc = prm + XO_PS_PRM_ICC + 34 * env;
  for (int g = 0; g < 22; g++) {
    int bin = g2b[g];
    int ii = iid[bin], ic = icc[bin];
    i16 c1 = sfac[steps + ii], c2 = sfac[steps - ii];
    i32 beta = ox_mul32x16_shl(ox_shl32(ox_mult16x16(alpha_tab[ic], ox_sub16(c1, c2)), 1), 0x5a82);
    i32 alpha = ox_shr32_dir_sat_limit((i32)alpha_tab[ic] << 16, 1);
    i16 bpa = ox_round16(ox_add_sat(beta, alpha)), bma = ox_round16(ox_sub_sat(beta, alpha));
    i32 ipa = ox_mul32x16(rescale, bpa), ima = ox_mul32x16(rescale, bma);
    i16 h11 = ox_mult16_shl(cos512(ipa, trig), c2), h12 = ox_mult16_shl(cos512(ima, trig), c1);
    i16 h21 = ox_mult16_shl(sin512(ipa, trig), c2), h22 = ox_mult16_shl(sin512(ima, trig), c1);
    d11[2 * g] = ox_mult16_shl(inv_len, ox_sub16(h11, h11v[2 * g]));
    d11[2 * g + 1] = ox_mult16_shl(inv_len, ox_sub16(h12, h11v[2 * g + 1]));
    d21[2 * g] = ox_mult16_shl(inv_len, ox_sub16(h21, h21v[2 * g]));
    d21[2 * g + 1] = ox_mult16_shl(inv_len, ox_sub16(h22, h21v[2 * g + 1]));
    H11[2 * g] = h11v[2 * g]; H11[2 * g + 1] = h11v[2 * g + 1];
    H21[2 * g] = h21v[2 * g]; H21[2 * g + 1] = h21v[2 * g + 1];
    h11v[2 * g] = h11; h11v[2 * g + 1] = h12;
    h21v[2 * g] = h21; h21v[2 * g + 1] = h22;
  }
}

/* decoder/ixheaacd_ps_dec.c:856-991.  left/right: QMF rows (re|im), modified in place */
static void apply_rot(i16 *ps, i32 *left, i32 *right, i32 *hyb, const uint8_t *ps_rom, int as_built) {
  i16 *idx = ps + XO_PS_ST_IDX, *hv = ps + XO_PS_ST_HVEC;
  i16 *H11 = hv + 96, *H21 = hv + 144, *d11 = hv + 192, *d21 = hv + 240;
  const i16 *borders = T16P(XO_PSROM_BORDERS_GROUP), *resol = T16P(XO_PSROM_HYB_RESOL);
  i32 *lre = hyb, *lim = hyb + 16, *rre = hyb + 32, *rim = hyb + 48;
  const int usb = idx[XO_PS_IDX_USB];
  for (int g = 0; g < 22; g++) {
    H11[2 * g] = ox_add16(H11[2 * g], d11[2 * g]);
    H11[2 * g + 1] = ox_add16(H11[2 * g + 1], d11[2 * g + 1]);
    H21[2 * g] = ox_add16(H21[2 * g], d21[2 * g]);
    H21[2 * g + 1] = ox_add16(H21[2 * g + 1], d21[2 * g + 1]);
  }
  for (int s = 0; s < 10; s++) {
    i32 a = ox_add_sat(M(lre[s], H11[2 * s]), M(rre[s], H21[2 * s]));
    i32 b = ox_add_sat(M(lim[s], H11[2 * s]), M(rim[s], H21[2 * s]));
    i32 c = ox_add_sat(M(lre[s], H11[2 * s + 1]), M(rre[s], H21[2 * s + 1]));
    i32 d = ox_add_sat(M(lim[s], H11[2 * s + 1]), M(rim[s], H21[2 * s + 1]));
    lre[s] = ox_shl32(a, 2); lim[s] = ox_shl32(b, 2); rre[s] = ox_shl32(c, 2); rim[s] = ox_shl32(d, 2);
  }
  i16 h[64][4];
  memset(h, 0, sizeof(h));
  /* :929-944 fills a local WORD16 H11_H12[256] = {0} through WORD32 pointers and :965-990 reads it back as WORD16.
   * That type punning is undefined in ISO C; the reference's own x86-64 build (gcc 13, -O3, no -fno-strict-aliasing,
   * cmake/utils.cmake:17) reads the array as still zero, so bands 3..usb-1 of both channels leave the rotation as 0.
   * as_built = 1 reproduces that build (what the tapped golden records contain); as_built = 0 is the source as
   * written (pinned against the same file compiled with -fno-strict-aliasing, oracle/_ref/libxaac_ref_nsa.so). */
  for (int g = 10; g < 22 && !as_built; g++) {
    int mxs = usb < borders[g + 1] ? usb : borders[g + 1];
    for (int s = borders[g]; s < mxs; s++) {
      h[s][0] = H11[2 * g]; h[s][1] = H11[2 * g + 1]; h[s][2] = H21[2 * g]; h[s][3] = H21[2 * g + 1];
    }
  }
  int o = 0, s;
  for (s = 0; s < 3; s++) {
    i32 a = lre[o], b = lim[o], c = rre[o], d = rim[o];
    int res = resol[s] < 6 ? resol[s] : 6;
    o++;
    for (int k = res - 2; k >= 0; k--, o++) {
      a = ox_add_sat(a, lre[o]); b = ox_add_sat(b, lim[o]); c = ox_add_sat(c, rre[o]); d = ox_add_sat(d, rim[o]);
    }
    left[s] = a; left[64 + s] = b; right[s] = c; right[64 + s] = d;
  }
  for (; s < usb; s++) {
    i32 a = ox_add_sat(M(left[s], h[s][0]), M(right[s], h[s][2]));
    i32 b = ox_add_sat(M(left[64 + s], h[s][0]), M(right[64 + s], h[s][2]));
    i32 c = ox_add_sat(M(left[s], h[s][1]), M(right[s], h[s][3]));
    i32 d = ox_add_sat(M(left[64 + s], h[s][1]), M(right[64 + s], h[s][3]));
    left[s] = ox_shl32(a, 2); left[64 + s] = ox_shl32(b, 2); right[s] = ox_shl32(c, 2); right[64 + s] = ox_shl32(d, 2);
  }
}

/* The per-slot PS work of the left synthesis call (decoder/ixheaacd_qmf_dec.c:1003-1030): ixheaacd_init_rot_env at
 * the PS envelope borders, ixheaacd_apply_ps (decoder/ixheaacd_thumb_ps_dec.c:69-99), then ixheaacd_shiftrountine
 * (generic:1610-1636) on the left row.  m: rows 0..37 already in the PS scale for rows < 32; right: 32 rows out. */
void xo_ps_apply_frame(const uint8_t *env_rom, const uint8_t *misc_rom, const uint8_t *ps_rom, const i16 *ps_prm,
                       i16 *ps, i32 *m, i32 *right, int usb, int shiftdelay_late, int common_shift, int as_built) {
  const i16 *inv_int = (const i16 *)(env_rom + XO_EROM_INV_INT), *trig = (const i16 *)misc_rom;
  i32 hyb[64];
  memset(hyb, 0, sizeof(hyb));
  int env = 0;
  for (int i = 0; i < 32; i++) {
    i32 *row = m + 128 * i, *rrow = right + 128 * i;
    if (env < 7 && i == ps_prm[XO_PS_PRM_BORDER + env]) {
      init_rot_env(ps, ps_prm, env, usb, ps_rom, inv_int, trig);
      env++;
    }
    int shiftdelay = i < 26 ? 0 : shiftdelay_late; /* thumb_ps_dec.c:78-80 */
    hybrid_analysis(m + 128 * (i + 6), hyb, (i32 *)(ps + XO_PS_ST_HYB), shiftdelay, ps_rom);
    decorrelation(ps, row, rrow, hyb, ps_rom);
    apply_rot(ps, row, rrow, hyb, ps_rom, as_built);
    if (common_shift) {
      for (int k = 0; k < 128; k++)
        row[k] = common_shift < 0 ? ox_shr32(row[k], -common_shift > 31 ? 31 : -common_shift)
                                  : ox_shl32_sat(row[k], common_shift);
    }
  }
}

/* The PS branch of ixheaacd_sbr_dec (decoder/ixheaacd_sbr_dec.c:1247-1272): ixheaacd_init_ps_scale, the left
 * ixheaacd_cplx_synt_qmffilt call with active = 1 (decoder/ixheaacd_qmf_dec.c:811-1129: pre-shifts, per-slot
 * ixheaacd_init_rot_env / ixheaacd_apply_ps, common shift, modulation, window) and the right call on the decorrelated
 * matrix.  m: matrix rows 0..37 (rows 0..31 are consumed). */
void xo_ps_synth_pair(const uint8_t *qrom, const uint8_t *env_rom, const uint8_t *misc_rom, const uint8_t *ps_rom,
                      const i16 *ps_prm, i16 *st, i16 *ps, i32 *m, i16 *out_l, i16 *out_r, int ch_out, int as_built) {
  i16 *sf = st + XO_SBR_ST_SF, *misc = st + XO_SBR_ST_MISC, *idx = ps + XO_PS_ST_IDX;
  i16 *sf_r = ps + XO_PS_ST_SF_R;
  static __thread i32 right[32 * 128];
  ps_init_scale(ps, sf, T16P(XO_PSROM_REV_DELAY));
  const int ps_scale = sf[XO_SF_PS], lsb = misc[XO_SBR_MISC_SYN_LSB], usb = misc[XO_SBR_MISC_SYN_USB];
  const int ov_lb_shift = ps_scale - sf[XO_SF_OV_LB], lb_shift = ps_scale - sf[XO_SF_LB];
  const int hb_shift = ps_scale - sf[XO_SF_HB], common_shift = (sf[XO_SF_ST_SYN] - ps_scale) - 8;
  if (ov_lb_shift == lb_shift) xo_adjust_scale_hq(m, 0, lsb, 0, 32, ov_lb_shift);
  else {
    xo_adjust_scale_hq(m, 0, lsb, 0, 6, ov_lb_shift);
    xo_adjust_scale_hq(m, 0, lsb, 6, 32, lb_shift);
  }
  xo_adjust_scale_hq(m, lsb, usb, 0, 32, hb_shift);
  xo_ps_apply_frame(env_rom, misc_rom, ps_rom, ps_prm, ps, m, right, usb, (i16)(sf[XO_SF_LB] - ps_scale), common_shift,
                    as_built);
  (void)idx;
  i32 sfv[4] = {sf[XO_SF_ST_SYN] - 8, sf[XO_SF_ST_SYN] - 8, sf[XO_SF_ST_SYN] - 8, sf[XO_SF_ST_SYN]};
  i32 off = st[XO_SBR_ST_SYN_POS], fpos = st[XO_SBR_ST_SYN_POS + 1];
  xo_synt_qmffilt_hq(qrom, m, st + XO_SBR_ST_SYN_STATES, &off, &fpos, sfv, lsb, usb, 6, out_l, ch_out);
  st[XO_SBR_ST_SYN_POS] = (i16)off;
  st[XO_SBR_ST_SYN_POS + 1] = (i16)fpos;
  sf_r[XO_SF_OV_LB] = sf_r[XO_SF_LB] = sf_r[XO_SF_HB] = (i16)ps_scale; /* sbr_dec.c:1259-1262 */
  i32 sfr[4] = {ps_scale, ps_scale, ps_scale, sf_r[XO_SF_ST_SYN]};
  off = ps[XO_PS_ST_SYN_POS_R];
  fpos = ps[XO_PS_ST_SYN_POS_R + 1];
  xo_synt_qmffilt_hq(qrom, right, ps + XO_PS_ST_SYN_STATES_R, &off, &fpos, sfr, idx[XO_PS_IDX_LSB_R], idx[XO_PS_IDX_USB_R],
                     6, out_r, ch_out);
  ps[XO_PS_ST_SYN_POS_R] = (i16)off;
  ps[XO_PS_ST_SYN_POS_R + 1] = (i16)fpos;
}
