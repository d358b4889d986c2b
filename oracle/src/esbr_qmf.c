/*
 * oracle/src/esbr_qmf.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain-C restatement of the 64-band eSBR QMF synthesis bank of libxaac (first piece of the float eSBR
 * path): the per-slot core of ixheaacd_esbr_synthesis_filt_block (decoder/ixheaacd_sbr_dec.c:583-654, stereo_config_idx
 * <= 0, 64 synthesis channels): float -> WORD32 (x 64), ixheaacd_esbr_inv_modulation (decoder/ixheaacd_qmf_dec.c:733) =
 * ixheaacd_esbr_cos_sin_mod with ixheaacd_esbr_radix4bfly / ixheaacd_esbr_postradixcompute2
 * (decoder/generic/ixheaacd_qmf_dec_generic.c:1163-1461, 880-973, 975-1057), ixheaacd_shiftrountine_with_rnd_hq (:1704),
 * ixheaacd_esbr_qmfsyn64_winadd (:1544), WORD32 -> float (/ 65536).  The arithmetic inside is integer (WORD32 data,
 * WORD32 twiddles, WORD64 products), so the float output is bit-exact as well.  Same index structure as the WORD16-twiddle
 * bank in qmf.c; pinned against the compiled reference by tests/test_oracle_esbr.py.
 */
#include <string.h>
#include "fixmath.h"
#include "xaac_oracle.h"

#define E32(off) ((const i32 *)(erom + (off)))
static inline i32 padd(i32 a, i32 w1, i32 b, i32 w2) { /* (a w1 + b w2) >> 32, plain 64-bit add */
  return (i32)(((i64)a * w1 + (i64)b * w2) >> 32);
}
static inline i32 psub(i32 a, i32 w1, i32 b, i32 w2) { /* ixheaac_sub64_sat(a w1, b w2) >> 32 */
  const i64 x = (i64)a * w1, y = (i64)b * w2;
  i64 d;
  if (__builtin_sub_overflow(x, y, &d)) d = x < 0 ? INT64_MIN : INT64_MAX;
  return (i32)(d >> 32);
}
static inline i32 psubw(i32 a, i32 w1, i32 b, i32 w2) { /* plain 64-bit subtraction (radix-4 stage, generic:948-966) */
  return (i32)((i64)((uint64_t)((i64)a * w1) - (uint64_t)((i64)b * w2)) >> 32);
}

/* generic:880-973 */
static void radix4_stage32(const i32 *w, i32 *x, int groups, int span) {
  for (int g = 0; g < groups; g++)
    for (int i = 0; i < span; i++) {
      i32 *e0 = x + 2 * (g * 4 * span + i), *e1 = e0 + 2 * span, *e2 = e0 + 4 * span, *e3 = e0 + 6 * span;
      const i32 *tw = w + 6 * i;
      const i32 si1 = tw[0], co1 = tw[1], si2 = tw[2], co2 = tw[3], si3 = tw[4], co3 = tw[5];
      i32 xh0 = ox_add_sat(e0[0], e2[0]), xl0 = ox_sub_sat(e0[0], e2[0]);
      i32 xh20 = ox_add_sat(e1[0], e3[0]), xl20 = ox_sub_sat(e1[0], e3[0]);
      i32 xh1 = ox_add_sat(e0[1], e2[1]), xl1 = ox_sub_sat(e0[1], e2[1]);
      i32 xh21 = ox_add_sat(e1[1], e3[1]), xl21 = ox_sub_sat(e1[1], e3[1]);
      i32 xt0 = ox_sub_sat(xh0, xh20), yt0 = ox_sub_sat(xh1, xh21);
      i32 xt1 = ox_add_sat(xl0, xl21), xt2 = ox_sub_sat(xl0, xl21);
      i32 yt2 = ox_add_sat(xl1, xl20), yt1 = ox_sub_sat(xl1, xl20);
      e0[0] = ox_add_sat(xh0, xh20);
      e0[1] = ox_add_sat(xh1, xh21);
      e3[0] = ox_shl1(padd(yt2, si3, xt2, co3));
      e3[1] = ox_shl1(psubw(yt2, co3, xt2, si3));
      e2[0] = ox_shl1(padd(yt0, si2, xt0, co2));
      e2[1] = ox_shl1(psubw(yt0, co2, xt0, si2));
      e1[0] = ox_shl1(padd(yt1, si1, xt1, co1));
      e1[1] = ox_shl1(psubw(yt1, co1, xt1, si1));
    }
}

/* generic:975-1057 with dig_rev_table2_32 = {0, 64, 16, 80} */
static void post_radix2_32e(i32 *y, const i32 *x) {
  static const int dr[4] = {0, 64, 16, 80};
  i32 *y0 = y, *y1 = y + 8, *y2 = y + 32, *y3 = y + 40;
  for (int blk = 0; blk < 4; blk++) {
    const int h2 = dr[blk] >> 2;
    const i32 *a = x + (blk >> 1) * 32 + (blk & 1) * 8;
    for (int half = 0; half < 2; half++) {
      const i32 *c = a + 16 * half;
      const int o = h2 + 2 * half;
      y0[o] = ox_add_sat(c[0], c[2]); y0[o + 1] = ox_add_sat(c[1], c[3]);
      y2[o] = ox_sub_sat(c[0], c[2]); y2[o + 1] = ox_sub_sat(c[1], c[3]);
      y1[o] = ox_add_sat(c[4], c[6]); y1[o + 1] = ox_add_sat(c[5], c[7]);
      y3[o] = ox_sub_sat(c[4], c[6]); y3[o + 1] = ox_sub_sat(c[5], c[7]);
    }
  }
}

/* generic:1059-1161 with dig_rev_table4_16 = {0, 16}: final radix-4 (no twiddles) of the 16-point FFT */
static void post_radix4_16e(i32 *y, const i32 *x) {
  i32 *y0 = y, *y1 = y + 8, *y2 = y + 16, *y3 = y + 24;
  for (int k = 0; k < 2; k++) {
    const int h2 = (16 * k) >> 2;
    for (int half = 0; half < 2; half++) {
      const i32 *c = x + 16 * k + 8 * half;
      const int o = h2 + 2 * half;
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

/* generic:1163-1461 for no_channels = 64 (M = 32, synthesis) or 32 (M = 16, analysis): sb[0..2M-1] and sb[64..64+2M-1] */
static void esbr_cos_sin_mod(const uint8_t *erom, i32 *sb, int no_channels) {
  const int M = no_channels >> 1, N = 2 * M, H = M >> 1;
  const i32 *tw = E32(no_channels == 64 ? XO_EROM2_SINCOS_L64 : XO_EROM2_SINCOS_L32);
  const i32 *alt = E32(no_channels == 64 ? XO_EROM2_ALTSIN_L64 : XO_EROM2_ALTSIN_L32);
  i32 t[128];
  i32 *s1 = sb, *s2 = sb + 64, *t1 = t, *t2 = t + 64;
  for (int n = 0; n < M; n++) { /* :1200-1295 */
    const i32 wim = tw[2 * n], wre = tw[2 * n + 1];
    const i32 a = s1[n], b = s1[N - 1 - n], c = s2[n], d = s2[N - 1 - n];
    if (!(n & 1)) {
      const int j = n >> 1;
      t1[2 * j] = padd(a, wre, b, wim);
      t1[2 * j + 1] = psub(b, wre, a, wim);
      t2[2 * j] = psub(d, wim, c, wre);
      t2[2 * j + 1] = padd(c, wim, d, wre);
    } else {
      const int j = (n - 1) >> 1;
      t1[N - 1 - 2 * j] = psub(a, wre, b, wim);
      t1[N - 2 - 2 * j] = padd(b, wre, a, wim);
      t2[N - 1 - 2 * j] = padd(d, wim, c, wre);
      t2[N - 2 - 2 * j] = psub(c, wim, d, wre);
    }
  }
  for (int h = 0; h < 2; h++) { /* :1297-1314 */
    if (M == 32) {
      radix4_stage32(E32(XO_EROM2_W32), t + 64 * h, 1, 8);
      radix4_stage32(E32(XO_EROM2_W32) + 48, t + 64 * h, 4, 2);
      post_radix2_32e(sb + 64 * h, t + 64 * h);
    } else {
      radix4_stage32(E32(XO_EROM2_W16), t + 64 * h, 1, 4);
      post_radix4_16e(sb + 64 * h, t + 64 * h);
    }
  }
  i32 f1[64], f2[64]; /* post-twiddle :1365-1460, restated out of place */
  memcpy(f1, s1, sizeof(i32) * N);
  memcpy(f2, s2, sizeof(i32) * N);
  s1[0] = f1[0] >> 1;
  s1[N - 1] = ox_neg_sat(f1[1] >> 1);
  s2[N - 1] = ox_neg_sat(f2[0] >> 1);
  s2[0] = f2[1] >> 1;
  for (int u = 0; u < H; u++) {
    const i32 wim = alt[2 * u], wre = alt[2 * u + 1];
    i32 re = f1[N - 1 - 2 * u], im = f1[N - 2 - 2 * u];
    s1[N - 2 - 2 * u] = padd(re, wre, im, wim);
    s1[1 + 2 * u] = psub(im, wre, re, wim);
    re = f2[N - 1 - 2 * u];
    im = f2[N - 2 - 2 * u];
    s2[1 + 2 * u] = ox_neg_sat(padd(re, wre, im, wim));
    s2[N - 2 - 2 * u] = psub(re, wim, im, wre);
    if (u + 1 < H) {
      i32 fim = f1[2 + 2 * u], fre = f1[3 + 2 * u];
      s1[2 + 2 * u] = padd(fre, wim, fim, wre);
      s1[N - 3 - 2 * u] = psub(fim, wim, fre, wre);
      fim = f2[2 + 2 * u];
      fre = f2[3 + 2 * u];
      s2[N - 3 - 2 * u] = ox_neg_sat(padd(fre, wim, fim, wre));
      s2[2 + 2 * u] = psub(fre, wre, fim, wim);
    }
  }
}
static void esbr_cos_sin_mod64(const uint8_t *erom, i32 *sb) { esbr_cos_sin_mod(erom, sb, 64); }

/* The per-slot core of ixheaacd_esbr_synthesis_filt_block for 32 time slots.
 *   qmf   [32][128] float: re[64] | im[64] per slot (qmf_buf_real[i][k], qmf_buf_imag[i][k])
 *   fs    [1280] WORD32 filter_states_32 (in/out); *off = ixheaacd_drc_offset, *fpos = filter_pos_syn_32 - esbr_qmf_c (in/out)
 *   out   [2048] float time samples */
void xo_esbr_synth64(const uint8_t *erom, const float *qmf, i32 *fs, i32 *off_io, i32 *fpos_io, float *out) {
  const i32 *qc = E32(XO_EROM2_QMF_C);
  int off = *off_io, fpos = *fpos_io;
  for (int i = 0; i < 32; i++) {
    i32 buf[128];
    for (int k = 0; k < 64; k++) { /* sbr_dec.c:584-587: C cast = truncation toward zero */
      buf[k] = (i32)(qmf[128 * i + k] * 64);
      buf[64 + k] = (i32)(qmf[128 * i + 64 + k] * 64);
    }
    esbr_cos_sin_mod64(erom, buf);
    i32 *st = fs + off; /* generic:1704-1734 with len = 64, shift = 6 */
    for (int j = 0; j < 32; j++) {
      const i32 r1 = buf[j], i1 = buf[64 + j], r2 = buf[63 - j], i2 = buf[64 + 63 - j];
      st[64 + 63 - j] = ox_shl32_sat(ox_add_sat(i1, r1), 6);
      st[63 - j] = ox_shl32_sat(ox_sub_sat(i2, r2), 6);
      st[j] = ox_shl32_sat(ox_sub_sat(i1, r1), 6);
      st[64 + j] = ox_shl32_sat(ox_add_sat(i2, r2), 6);
    }
    const i32 *fp1 = fs + ((i & 1) ? 64 : 0), *fp2 = fs + ((i & 1) ? 0 : 64);
    const i32 *c = qc + fpos;
    for (int k = 0; k < 64; k++) { /* generic:1544-1575 */
      uint64_t acc = 0;
      for (int j = 0; j < 5; j++) acc += (uint64_t)((i64)fp1[256 * j + k] * c[k + 128 * j]);
      for (int j = 0; j < 5; j++) acc += (uint64_t)((i64)fp2[128 + 256 * j + k] * c[k + 64 + 128 * j]);
      out[64 * i + k] = (float)(i32)((i64)acc >> 31) / 65536.0f;
    }
    off -= 128;
    if (off < 0) off += 1280;
    fpos += 64;
    if (fpos == 640) fpos = 0;
  }
  *off_io = off;
  *fpos_io = fpos;
}
void xo_esbr_synth64_batch(const uint8_t *erom, const float *qmf, i32 *fs, i32 *pos, float *out, int n) {
  for (int u = 0; u < n; u++)
    xo_esbr_synth64(erom, qmf + (size_t)u * 4096, fs + (size_t)u * 1280, pos + 2 * u, pos + 2 * u + 1, out + (size_t)u * 2048);
}

/* ixheaacd_esbr_analysis_filt_block (decoder/ixheaacd_sbr_dec.c:185-295) for 32 analysis channels and 32 time slots:
 * float core samples -> WORD32 (x 2^15), ixheaacd_esbr_qmfanal32_winadd (decoder/ixheaacd_qmf_dec.c:537-640, WORD64
 * accumulation), ixheaacd_esbr_fwd_modulation (generic:1463-1506: >> 4, fold, cos_sin_mod M = 16, t_cos rotation),
 * WORD32 -> float (x 1/256).
 *   time_in [1024] float; states [320] WORD32 anal_filter_states_32 (in/out); *pos = state_new_samples_pos_low_32 -
 *   anal_filter_states_32, *fpos = filter_pos_32 - esbr_qmf_c (in/out); qmf [32][128] float: re at +0..31, im at +64..95 */
void xo_esbr_anal32(const uint8_t *erom, const float *time_in, i32 *states, i32 *pos_io, i32 *fpos_io, float *qmf) {
  const i32 *qc = E32(XO_EROM2_QMF_C), *tcos = E32(XO_EROM2_TCOS_L32);
  int pos = *pos_io, f1 = *fpos_io, f2 = f1 + 64;
  for (int i = 0; i < 32; i++) {
    i32 buf[64], sb[128];
    for (int z = 0; z < 32; z++) states[pos + 31 - z] = (i32)(time_in[32 * i + z] * (1 << 15));
    const i32 *fp1 = states + ((i & 1) ? 32 : 0), *fp2 = states + ((i & 1) ? 0 : 32);
    for (int n = 0; n < 32; n++) {
      uint64_t a = 0, b = 0;
      for (int j = 0; j < 5; j++) {
        a += (uint64_t)((i64)fp1[n + 64 * j] * qc[f1 + 2 * (n + 64 * j)]);
        b += (uint64_t)((i64)fp2[n + 64 * j] * qc[f2 + 2 * (n + 64 * j)]);
      }
      buf[n] = (i32)((i64)a >> 31);
      buf[n + 32] = (i32)((i64)b >> 31);
    }
    pos -= 32;
    if (pos < 0) pos = 288;
    {
      const int n1 = f2 + 64, n2 = f1 + 64;
      f1 = n1;
      f2 = n2;
      if (f2 > 640) { f1 = 0; f2 = 64; }
    }
    memset(sb, 0, sizeof(sb));
    for (int k = 0; k < 32; k++) { /* generic:1475-1482 */
      const i32 t1 = ox_shr32(buf[k], 4), t2 = ox_shr32(buf[63 - k], 4);
      sb[k] = ox_sub_sat(t1, t2);
      sb[64 + k] = ox_add_sat(t1, t2);
    }
    esbr_cos_sin_mod(erom, sb, 32);
    for (int k = 0; k < 32; k++) { /* generic:1490-1505, usb - lsb = 32 */
      const i32 ch = tcos[2 * k], sh = tcos[2 * k + 1], re = sb[k], im = sb[64 + k];
      const i32 r2 = (i32)(((i64)re * ch + (i64)im * sh) >> 31);
      const i64 x = (i64)im * ch, y = (i64)re * sh;
      i64 d;
      if (__builtin_sub_overflow(x, y, &d)) d = x < 0 ? INT64_MIN : INT64_MAX;
      qmf[128 * i + k] = (float)r2 * (1.0f / 256.0f);
      qmf[128 * i + 64 + k] = (float)(i32)(d >> 31) * (1.0f / 256.0f);
    }
  }
  *pos_io = pos;
  *fpos_io = f1;
}
void xo_esbr_anal32_batch(const uint8_t *erom, const float *time_in, i32 *states, i32 *pos, float *qmf, int n) {
  for (int u = 0; u < n; u++)
    xo_esbr_anal32(erom, time_in + (size_t)u * 1024, states + (size_t)u * 320, pos + 2 * u, pos + 2 * u + 1, qmf + (size_t)u * 4096);
}


/* ---- hand-overs around the eSBR stage ----------------------------------------------------------------- */
/* USAC core -> eSBR: time_sample_vector = (FLOAT32)output_data_ptr * (FLOAT32)ONE_BY_TWO_POW_15 (ixheaacd_ext_ch_ele.c:1040) */
void xo_esbr_core_to_float(const int32_t *core, float *out, int n) {
  for (int k = 0; k < n; k++) out[k] = (float)((float)core[k] * (float)(0.000030517578125));
}
/* legacy core -> eSBR: (FLOAT32)time_data[ch_fac * i + ch], no scaling (ixheaacd_api.c:3384-3437) */
void xo_esbr_pcm16_to_float(const int16_t *pcm, int ch_fac, int ch, float *out, int n) {
  for (int i = 0; i < n; i++) out[i] = (float)pcm[ch_fac * i + ch];
}
/* ixheaacd_samples_sat, pcmsize 16 (ixheaacd_decode_main.c:82-104): clamp, then the C cast (truncation), interleaved */
void xo_samples_sat16(const float *in, int ch_fac, int ch, int16_t *pcm, int n) {
  for (int i = 0; i < n; i++) {
    float v = in[i];
    if (v > 32767.0f)
      v = 32767.0f;
    else if (v < -32768.0f)
      v = -32768.0f;
    pcm[ch_fac * i + ch] = (int16_t)v;
  }
}
