/*
 * oracle/src/esbr_hbe.c — TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Plain-C restatement of the QMF-domain harmonic transposer of the float eSBR decoder:
 *   ixheaacd_qmf_hbe_apply            decoder/ixheaacd_hbe_trans.c:224-296
 *   ixheaacd_real_synth_filt          decoder/ixheaacd_esbr_polyphase.c:157-274
 *   ixheaacd_complex_anal_filt        decoder/ixheaacd_esbr_polyphase.c:48-155
 *   ixheaacd_hbe_post_anal_process    decoder/ixheaacd_hbe_trans.c:1549-1606 with prod2/3/4 (:757-1059),
 *                                     xprod2/3/4 (:1061-1547), xprod_proc_3/4 (:371-755), norm_qmf_in_buf_2/4 (:298-369)
 *   ixheaac_real_synth_fft_p2/_p3, ixheaac_cmplx_anal_fft_p2/_p3   common/ixheaac_esbr_fft.c:42, 1084, 537, 1148
 * for the 2:1 system (32 QMF columns per call).  Every float operation is one IEEE rounding in the reference's order (the
 * reference build has no FMA and evaluates float expressions in float, double ones in double).  Pinned against the
 * compiled reference function through oracle/ref_shim_hbe.c and against records tapped from real decodes.
 */
#include <math.h>
#include <string.h>
#include "xaac_oracle.h"

typedef struct { float r, i; } cf;

/* ---- radix-4 FFT of common/ixheaac_esbr_fft.c: the stages after the first are shared by both transforms (:95-535) ---- */
static void rot_a(cf *x, float wc, float ws) { /* :216-218 */
  float t = x->r * wc + x->i * ws;
  x->i = -(x->r * ws) + x->i * wc;
  x->r = t;
}
static void rot_b(cf *x, float w3, float w6) { /* :275-277 */
  float t = x->r * w6 - x->i * w3;
  x->i = x->r * w3 + x->i * w6;
  x->r = t;
}
static void rot_c(cf *x, float w3, float w6) { /* :431-433 */
  float t = -(x->r * w3) - x->i * w6;
  x->i = -(x->r * w6) + x->i * w3;
  x->r = t;
}
/* the radix-4 butterfly and its store order: y0 = x0, y1 = x2, y2 = x1, y3 = (x3i, x3r); alt = last twiddle segment (:438-445) */
static void bfly4(cf x0, cf x1, cf x2, cf x3, int alt, float *d, int st) {
  x0.r = x0.r + x2.r;
  x0.i = x0.i + x2.i;
  x2.r = x0.r - (x2.r * 2);
  x2.i = x0.i - (x2.i * 2);
  x1.r = x1.r + x3.r;
  if (!alt) {
    x1.i = x1.i + x3.i;
    x3.r = x1.r - (x3.r * 2);
    x3.i = x1.i - (x3.i * 2);
  } else {
    x1.i = x1.i - x3.i;
    x3.r = x1.r - (x3.r * 2);
    x3.i = x1.i + (x3.i * 2);
  }
  x0.r = x0.r + x1.r;
  x0.i = x0.i + x1.i;
  x1.r = x0.r - (x1.r * 2);
  x1.i = x0.i - (x1.i * 2);
  x2.r = x2.r - x3.i;
  x2.i = x2.i + x3.r;
  x3.i = x2.r + (x3.i * 2);
  x3.r = x2.i - (x3.r * 2);
  d[0] = x0.r; d[1] = x0.i;
  d[st] = x2.r; d[st + 1] = x2.i;
  d[2 * st] = x1.r; d[2 * st + 1] = x1.i;
  d[3 * st] = x3.i; d[3 * st + 1] = x3.r;
}

static unsigned dig_rev(unsigned v, int m) { /* :29-36 */
  v = ((v & 0x33333333u) << 2) | ((v & ~0x33333333u) >> 2);
  v = ((v & 0x0F0F0F0Fu) << 4) | ((v & ~0x0F0F0F0Fu) >> 4);
  v = ((v & 0x00FF00FFu) << 8) | ((v & ~0x00FF00FFu) >> 8);
  return v >> m;
}
static int ilog2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }

static void fft_rest(const float *tw, float *y, int npoints) {
  const int lg = ilog2(npoints), not_power_4 = lg & 1, n_stages = lg >> 1;
  int del = 4, nodespacing = 64, in_loop_cnt = npoints >> 4;
  for (int s = n_stages - 1; s > 0; s--) {
    const int nd = nodespacing * del;
    const int sec = nd / 4 + nd / 8 - nd / 16 + nd / 32 - nd / 64 + nd / 128 - nd / 256;
    for (int jj = 0; jj < del; jj++) {
      const int j = jj * nodespacing;
      for (int k = 0; k < in_loop_cnt; k++) {
        float *d = y + 2 * jj + k * 8 * del;
        cf x0 = {d[0], d[1]}, x1 = {d[2 * del], d[2 * del + 1]}, x2 = {d[4 * del], d[4 * del + 1]},
           x3 = {d[6 * del], d[6 * del + 1]};
        int alt = 0;
        if (jj > 0) {
          rot_a(&x1, tw[j], tw[j + 257]);
          if (j <= sec) {
            rot_a(&x2, tw[2 * j], tw[2 * j + 257]);
            rot_a(&x3, tw[3 * j], tw[3 * j + 257]);
          } else if (j <= (nd >> 1)) {
            rot_a(&x2, tw[2 * j], tw[2 * j + 257]);
            rot_b(&x3, tw[3 * j - 256], tw[3 * j + 1]);
          } else if (j <= sec * 2) {
            rot_b(&x2, tw[2 * j - 256], tw[2 * j + 1]);
            rot_b(&x3, tw[3 * j - 256], tw[3 * j + 1]);
          } else {
            rot_b(&x2, tw[2 * j - 256], tw[2 * j + 1]);
            rot_c(&x3, tw[3 * j - 512], tw[3 * j - 512 + 257]);
            alt = 1;
          }
        }
        bfly4(x0, x1, x2, x3, alt, d, 2 * del);
      }
    }
    nodespacing >>= 2;
    del <<= 2;
    in_loop_cnt >>= 2;
  }
  if (not_power_4) { /* :480-533 */
    nodespacing <<= 1;
    for (int t = 0; t < del; t++) {
      const int q = t < del / 2 ? t : t - del / 2;
      const float w1 = tw[q * nodespacing], w4 = tw[q * nodespacing + 257];
      float *p = y + 2 * t;
      cf x0 = {p[0], p[1]}, x1 = {p[2 * del], p[2 * del + 1]};
      if (t < del / 2) rot_a(&x1, w1, w4);
      else rot_b(&x1, w1, w4);   /* tmp = x1r*W4 - x1i*W1; x1i = x1r*W1 + x1i*W4 */
      p[2 * del] = x0.r - x1.r;
      p[2 * del + 1] = x0.i - x1.i;
      p[0] = x0.r + x1.r;
      p[1] = x0.i + x1.i;
    }
  }
}

/* ixheaac_real_synth_fft_p2 (:42-535): npoints real inputs -> npoints complex outputs */
static void real_synth_fft_p2(const float *tw, const float *x, float *y, int npoints) {
  const int lg = ilog2(npoints), not_power_4 = lg & 1;
  const int shift = (31 - 1 - lg) + 1 - 16; /* ixheaac_norm32(npoints) + 1 - 16 */
  for (int i = 0; i < npoints; i += 4) {
    int h2 = (int)dig_rev((unsigned)i, shift);
    if (not_power_4) h2 = (h2 + 1) & ~1;
    const float *inp = x + (h2 >> 1);
    float x0r = inp[0], x1r = inp[npoints >> 2], x2r = inp[2 * (npoints >> 2)], x3r = inp[3 * (npoints >> 2)];
    x0r = x0r + x2r;
    x2r = x0r - (x2r * 2);
    x1r = x1r + x3r;
    x3r = x1r - (x3r * 2);
    x0r = x0r + x1r;
    x1r = x0r - (x1r * 2);
    float *o = y + 2 * i;
    o[0] = x0r; o[1] = 0; o[2] = x2r; o[3] = x3r; o[4] = x1r; o[5] = 0; o[6] = x2r; o[7] = -x3r;
  }
  fft_rest(tw, y, npoints);
}
/* ixheaac_cmplx_anal_fft_p2 (:537-1046) */
static void cmplx_anal_fft_p2(const float *tw, const float *x, float *y, int npoints) {
  const int lg = ilog2(npoints), not_power_4 = lg & 1;
  const int shift = (31 - 1 - lg) + 1 - 16;
  for (int i = 0; i < npoints; i += 4) {
    int h2 = (int)dig_rev((unsigned)i, shift);
    if (not_power_4) h2 = (h2 + 1) & ~1;
    const float *inp = x + h2;
    const int st = npoints >> 1;
    cf x0 = {inp[0], inp[1]}, x1 = {inp[st], inp[st + 1]}, x2 = {inp[2 * st], inp[2 * st + 1]},
       x3 = {inp[3 * st], inp[3 * st + 1]};
    bfly4(x0, x1, x2, x3, 0, y + 2 * i, 2);
  }
  fft_rest(tw, y, npoints);
}
/* ixheaac_aac_ld_dec_fft_3_float (:1048-1082) */
static void fft3(const float *inp, float *op) {
  const float sinmu = -0.866025403784439f;
  float temp_real = inp[0] + inp[2], temp_imag = inp[1] + inp[3];
  float add_r = inp[2] + inp[4], add_i = inp[3] + inp[5];
  float sub_r = inp[2] - inp[4], sub_i = inp[3] - inp[5];
  float p1 = add_r / 2.0f, p4 = add_i / 2.0f, p2 = sub_i * sinmu, p3 = sub_r * sinmu;
  float temp = inp[0] - p1;
  op[0] = temp_real + inp[4];
  op[1] = temp_imag + inp[5];
  op[2] = temp + p2;
  op[3] = (inp[1] - p3) - p4;
  op[4] = temp - p2;
  op[5] = (inp[1] + p3) - p4;
}
static void tw3(float *x, const float *wr, int n) { /* :1110-1129 / :1171-1192 */
  x += 2;
  for (int i = 0; i < n; i++) {
    for (int q = 0; q < 2; q++) {
      float t = x[0] * wr[0] + x[1] * wr[1];
      x[1] = -x[0] * wr[1] + x[1] * wr[0];
      x[0] = t;
      wr += 2;
      x += 2;
    }
    x += 2;
  }
}
/* ixheaac_real_synth_fft_p3 (:1084-1146), npoints = 24 */
static void real_synth_fft_p3(const float *rom, const float *x_in, float *x_out) {
  float x_3[8], y_3[16], y[48], x[48];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 8; j++) x_3[j] = x_in[3 * j + i];
    real_synth_fft_p2(rom + XO_HROM_FFTTW, x_3, y_3, 8);
    for (int j = 0; j < 16; j += 2) {
      x[3 * j + 2 * i] = y_3[j];
      x[3 * j + 2 * i + 1] = y_3[j + 1];
    }
  }
  tw3(x, rom + XO_HROM_TW24, 8);
  for (int i = 0; i < 8; i++) fft3(x + 6 * i, y + 6 * i);
  const float *py = y;
  for (int i = 0; i < 16; i += 2) {
    x_out[i] = *py++; x_out[i + 1] = *py++;
    x_out[16 + i] = *py++; x_out[16 + i + 1] = *py++;
    x_out[32 + i] = *py++; x_out[32 + i + 1] = *py++;
  }
}
/* ixheaac_cmplx_anal_fft_p3 (:1148-1209), npoints = 48; works in place on x_in like the reference */
static void cmplx_anal_fft_p3(const float *rom, float *x_in, float *x_out) {
  float x_3[32], y_3[32], y[96];
  for (int i = 0; i < 6; i += 2) {
    for (int j = 0; j < 32; j += 2) {
      x_3[j] = x_in[3 * j + i];
      x_3[j + 1] = x_in[3 * j + i + 1];
    }
    cmplx_anal_fft_p2(rom + XO_HROM_FFTTW, x_3, y_3, 16);
    for (int j = 0; j < 32; j += 2) {
      x_in[3 * j + i] = y_3[j];
      x_in[3 * j + i + 1] = y_3[j + 1];
    }
  }
  tw3(x_in, rom + XO_HROM_TW48, 16);
  for (int i = 0; i < 16; i++) fft3(x_in + 6 * i, y + 6 * i);
  const float *py = y;
  for (int i = 0; i < 32; i += 2) {
    x_out[i] = *py++; x_out[i + 1] = *py++;
    x_out[32 + i] = *py++; x_out[32 + i + 1] = *py++;
    x_out[64 + i] = *py++; x_out[64 + i + 1] = *py++;
  }
}

static int win_off(int len) { /* ixheaacd_map_prot_filter, hbe_trans.c:70-101 */
  switch (len) {
    case 4: return 0;
    case 8: return 40;
    case 12: return 120;
    case 16: return 240;
    case 20: return 400;
    case 24: return 600;
    case 32: return 840;
    case 40: return 1160;
  }
  return 0;
}
static int syncos_off(int s) { return s == 4 ? 0 : s == 8 ? 16 : s == 12 ? 48 : 96; }
static int anacs_off(int s) { return s == 4 ? 0 : s == 8 ? 32 : s == 12 ? 96 : 192; }

/* working set of one transposer instance (the reference keeps these in ia_esbr_hbe_txposer_struct / persistent memory;
 * qin is flat because the cross-product search may index past a row end exactly as the reference's pointer arithmetic does) */
typedef struct {
  const float *rom;
  int S, k_start, start_band, end_band, max_stretch, xo[6];
  float inbuf[34 * 20 + 8];
  float synth_buf[400], analy_buf[400];
  float qin[34 * 128];   /* rows 0..31 (+ slack for the reference's out-of-row reads) */
  float norm[34 * 128];
  float qout[64 * 128];
} hbe_t;

static void real_synth_filt(hbe_t *h, const float *qre, const float *qim) { /* polyphase.c:157-274 */
  const int S = h->S;
  const float *ct = h->rom + XO_HROM_COSTRANS + h->k_start * 32;
  const float *sct = h->rom + XO_HROM_SYNCOS + syncos_off(S);
  const float *win = h->rom + XO_HROM_WIN + win_off(S);
  float *buffer = h->synth_buf;
  for (int idx = 0; idx < 32; idx++) {
    float loc[64], so[128], g[400], w[400];
    float *out_buf = h->inbuf + (idx + 1) * S;
    for (int k = 0; k < S; k++) {
      const int ki = h->k_start + k;
      loc[k] = (float)(ct[(k << 1) + 0] * qre[idx * 64 + ki] + ct[(k << 1) + 1] * qim[idx * 64 + ki]);
      loc[k + S] = 0;
    }
    for (int l = 20 * S - 1; l >= 2 * S; l--) buffer[l] = buffer[l - 2 * S];
    if (S == 20) { /* polyphase.c:203-229: direct-form modulation */
      const float *pt = h->rom + XO_HROM_SYN20;
      for (int l = 0; l < S + 1; l++) {
        float accu = 0.0f;
        for (int k = 0; k < S; k++) accu += loc[k] * pt[k];
        buffer[0 + l] = accu;
        buffer[S - l] = accu;
        pt += S;
      }
      for (int l = S + 1; l < 2 * S - S / 2; l++) {
        float accu = 0.0f;
        for (int k = 0; k < S; k++) accu += loc[k] * pt[k];
        buffer[0 + l] = accu;
        buffer[3 * S - l] = -accu;
        pt += S;
      }
      float accu = 0.0f;
      for (int k = 0; k < S; k++) accu += loc[k] * pt[k];
      buffer[3 * S >> 1] = accu;
    } else {
    if (S == 12) real_synth_fft_p3(h->rom, loc, so);
    else real_synth_fft_p2(h->rom + XO_HROM_FFTTW, loc, so, 2 * S);
    {
      const float *pu = so, *pc = sct;
      int kmax = S >> 1;
      float *syn = &buffer[kmax];
      kmax += S;
      for (int k = 0; k < kmax; k++) {
        float tmp = pu[0] * pc[0];
        tmp -= pu[1] * pc[1];
        pu += 2; pc += 2;
        *syn++ = tmp;
      }
      syn = &buffer[0];
      kmax -= S;
      for (int k = 0; k < kmax; k++) {
        float tmp = pu[0] * pc[0];
        tmp -= pu[1] * pc[1];
        pu += 2; pc += 2;
        *syn++ = tmp;
      }
    }
    }
    for (int i = 0; i < 5; i++) {
      memcpy(&g[(2 * i + 0) * S], &buffer[(4 * i + 0) * S], sizeof(float) * S);
      memcpy(&g[(2 * i + 1) * S], &buffer[(4 * i + 3) * S], sizeof(float) * S);
    }
    for (int k = 0; k < 10 * S; k++) w[k] = g[k] * win[k];
    for (int i = 0; i < S; i++) {
      float accu = 0.0f;
      for (int j = 0; j < 10; j++) accu = accu + w[S * j + i];
      out_buf[i] = accu;
    }
  }
}

static void complex_anal_filt(hbe_t *h) { /* polyphase.c:48-155, esbr_hq == 0 */
  const int S = h->S, A = 2 * S, N = 10 * A;
  const float *cs = h->rom + XO_HROM_ANACS + anacs_off(S);
  const float *win = h->rom + XO_HROM_WIN + win_off(A);
  float *x = h->analy_buf;
  for (int idx = 0; idx < 16; idx++) {
    float wo[400], u[160], u_in[256], u_out[256];
    const float *inp = h->inbuf + idx * 2 * S + 1;
    float *row = h->qin + (idx + 12) * 128;
    memset(row, 0, 128 * sizeof(float));
    float *ab = row + 4 * h->k_start;
    for (int i = N - 1; i >= A; i--) x[i] = x[i - A];
    for (int i = A - 1; i >= 0; i--) x[i] = inp[A - 1 - i];
    for (int i = 0; i < N; i++) wo[i] = x[i] * win[i];
    for (int i = 0; i < 2 * A; i++) {
      float accu = 0.0f;
      for (int j = 0; j < 5; j++) accu = accu + wo[i + j * 2 * A];
      u[i] = accu;
    }
    if (A == 40) { /* polyphase.c:109-130: direct-form modulation */
      const float *pt = h->rom + XO_HROM_ANA40;
      for (int i = 1; i < A; i++) {
        float t1 = u[i] + u[2 * A - i], t2 = u[i] - u[2 * A - i];
        u[i] = t1;
        u[2 * A - i] = t2;
      }
      for (int k = 0; k < A; k++) {
        float ar = u[A], ai = (k & 1) ? u[0] : -u[0];
        for (int l = 1; l < A; l++) {
          ar = ar + u[0 + l] * pt[2 * l + 0];
          ai = ai + u[2 * A - l] * pt[2 * l + 1];
        }
        pt += 2 * A;
        *ab++ = ar;
        *ab++ = ai;
      }
      continue;
    }
    for (int k = 0; k < 2 * A; k++) {
      u_in[2 * k] = cs[2 * k] * u[k];
      u_in[2 * k + 1] = cs[2 * k + 1] * u[k];
    }
    if (S == 12) cmplx_anal_fft_p3(h->rom, u_in, u_out);
    else cmplx_anal_fft_p2(h->rom + XO_HROM_FFTTW, u_in, u_out, 2 * A);
    const float *v = u_out;
    for (int k = 0; k < A / 2; k++) {
      ab[1] = -v[0];
      ab[0] = v[1];
      ab[3] = v[2];
      ab[2] = -v[3];
      ab += 4;
      v += 4;
    }
  }
}

/* ---- magnitude normalisations ---- */
static float mag2(float xr, float xi) { /* norm_qmf_in_buf_2 :351-358, xprod2 :1143-1149 */
  double base = 1e-17;
  float t = xr * xr;
  base = base + t;
  base = base + xi * xi;
  float m = (float)(1.0f / base);
  return (float)sqrt(sqrt(m));
}
static float mag3(float xr, float xi) { /* prod3 :832-835: cbrt(1.0f / (FLOAT32)base) */
  double base = 1e-17;
  double b1 = base + xr * xr;
  b1 = b1 + xi * xi;
  return (float)cbrt(1.0f / (float)b1);
}
static float mag4(float xr, float xi) { /* norm_qmf_in_buf_4 :312-322 */
  double base = 1e-17;
  float t = xr * xr;
  base = base + t;
  t = xi * xi;
  base = base + t;
  t = (float)sqrt(sqrt(base));
  float m = t * (float)(sqrt(t));
  return 1 / m;
}
static void norm_rows(hbe_t *h, int b0, int b1, int mode) { /* bands b0..b1 inclusive, rows 0..31 */
  if (b1 > 63) b1 = 63; /* the reference's band 64 lands in the next row's first cell and is never read */
  for (int b = b0; b <= b1; b++)
    for (int i = 0; i < 32; i++) {
      if (b < 0) continue;
      float xr = h->qin[i * 128 + 2 * b], xi = h->qin[i * 128 + 2 * b + 1];
      float m = mode == 2 ? mag2(xr, xi) : mag4(xr, xi);
      h->norm[i * 128 + 2 * b] = xr * m;
      h->norm[i * 128 + 2 * b + 1] = xi * m;
    }
}
#define QIN(row, fl) h->qin[(row) * 128 + (fl)]
static float fmin_(float a, float b) { return a < b ? a : b; } /* the reference's min() macro */

static void cpow_n(float *xr, float *xi, int n) { /* :521-527: n-1 complex multiplications by the original value */
  const float tr = *xr, ti = *xi;
  for (int q = 0; q < n - 1; q++) {
    float tmp = *xr;
    *xr = *xr * tr - *xi * ti;
    *xi = tmp * ti + *xi * tr;
  }
}

/* ixheaacd_hbe_xprod_proc_3 (:371-553) */
static void xprod_proc_3(hbe_t *h, int band, int col, float p, int pidx) {
  const int inp_band = 2 * band / 3;
  const int zr = col + 6;
  float mag_zero = QIN(zr, 2 * inp_band) * QIN(zr, 2 * inp_band) + QIN(zr, 2 * inp_band + 1) * QIN(zr, 2 * inp_band + 1);
  float max_mag = 0;
  int max_n1 = 0, max_n2 = 0, max_tr = 0;
  for (int tr = 1; tr < 3; tr++) {
    double temp_fac = (2.0f * band + 1 - tr * p) * 0.3333334;
    int n1 = (int)(temp_fac), n2 = (int)(temp_fac + p);
    float m1 = QIN(zr, 2 * n1) * QIN(zr, 2 * n1) + QIN(zr, 2 * n1 + 1) * QIN(zr, 2 * n1 + 1);
    float m2 = QIN(zr, 2 * n2) * QIN(zr, 2 * n2) + QIN(zr, 2 * n2 + 1) * QIN(zr, 2 * n2 + 1);
    float t = fmin_(m1, m2);
    if (t > max_mag) { max_mag = t; max_tr = tr; max_n1 = n1; max_n2 = n2; }
  }
  if (!(max_mag > mag_zero && max_n1 >= 0 && max_n2 < 64)) return;
  const float *ic = h->rom + XO_HROM_INTERP;
  float vyr[2], vyi[2], vor[2], voi[2], d1, d2, xzr, xzi;
  int mid = 3 - max_tr, na, nb; /* na: band of the zero-band factor, nb: band of the interpolated pair */
  if (max_tr == 1) { d1 = 0; d2 = 1.5f; na = max_n1; nb = max_n2; }
  else { d1 = 1.5f; d2 = 0; mid = max_tr; max_tr = 3 - max_tr; na = max_n2; nb = max_n1; }
  xzr = QIN(zr, 2 * na);
  xzi = QIN(zr, 2 * na + 1);
  {
    int idx = ((nb & 3) + 1) & 3;
    float cr0 = ic[2 * idx], ci0 = ic[2 * idx + 1], cr1 = cr0, ci1 = -ci0;
    vyr[1] = QIN(zr, 2 * nb);
    vyi[1] = QIN(zr, 2 * nb + 1);
    float tr_ = QIN(zr - 2, 2 * nb), ti_ = QIN(zr - 2, 2 * nb + 1);
    vyr[0] = cr1 * tr_ - ci1 * ti_;
    vyi[0] = ci1 * tr_ + cr1 * ti_;
    tr_ = QIN(zr - 1, 2 * nb); ti_ = QIN(zr - 1, 2 * nb + 1);
    vyr[0] += cr0 * tr_ - ci0 * ti_;
    vyi[0] += ci0 * tr_ + cr0 * ti_;
  }
  {
    double base = 1e-17;
    base = base + xzr * xzr;
    base = base + xzi * xzi;
    float m = (float)cbrt(1.0f / (float)base);
    xzr *= m; xzi *= m;
    for (int k = 0; k < 2; k++) {
      base = 1e-17;
      base = base + vyr[k] * vyr[k];
      base = base + vyi[k] * vyi[k];
      m = (float)cbrt(1.0f / (float)base);
      vyr[k] *= m; vyi[k] *= m;
    }
  }
  cpow_n(&xzr, &xzi, mid);
  for (int k = 0; k < 2; k++) cpow_n(&vyr[k], &vyi[k], max_tr);
  for (int k = 0; k < 2; k++) {
    vor[k] = vyr[k] * xzr - vyi[k] * xzi;
    voi[k] = vyr[k] * xzi + vyi[k] * xzr;
  }
  {
    float c = h->rom[XO_HROM_XP3 + (pidx << 1)], s = h->rom[XO_HROM_XP3 + (pidx << 1) + 1];
    if (d2 < d1) s = -s;
    float tr_ = vor[0], ti_ = voi[0];
    vor[0] = (float)(c * tr_ - s * ti_);
    voi[0] = (float)(c * ti_ + s * tr_);
  }
  for (int k = 0; k < 2; k++) {
    h->qout[(col * 2 + k + 5) * 128 + 2 * band] += (float)(1.8856f * vor[k]);
    h->qout[(col * 2 + k + 5) * 128 + 2 * band + 1] += (float)(1.8856f * voi[k]);
  }
}

/* ixheaacd_hbe_xprod_proc_4 (:555-755); n1 / n2 are float offsets inside a row (2 x band) */
static void xprod_proc_4(hbe_t *h, int band, int col, float p, int pidx) {
  const int inp_band = band >> 1;
  const int zr = col + 6;
  float mag_zero = QIN(zr, 2 * inp_band) * QIN(zr, 2 * inp_band) + QIN(zr, 2 * inp_band + 1) * QIN(zr, 2 * inp_band + 1);
  float max_mag = 0;
  int max_n1 = 0, max_n2 = 0, max_tr = 0;
  for (int tr = 1; tr < 4; tr++) {
    double temp_fac = (2.0 * band + 1 - tr * p) * 0.25;
    int n1 = ((int)(temp_fac)) << 1, n2 = ((int)(temp_fac + p)) << 1;
    float m1 = QIN(zr, n1) * QIN(zr, n1) + QIN(zr, n1 + 1) * QIN(zr, n1 + 1);
    float m2 = QIN(zr, n2) * QIN(zr, n2) + QIN(zr, n2 + 1) * QIN(zr, n2 + 1);
    float t = fmin_(m1, m2);
    if (t > max_mag) { max_mag = t; max_tr = tr; max_n1 = n1; max_n2 = n2; }
  }
  if (!(max_mag > mag_zero && max_n1 >= 0 && max_n2 < 128)) return;
  float vyr[2], vyi[2], vor[2], voi[2], d1, d2, xzr, xzi;
  int mid = 4 - max_tr;
  if (max_tr == 1) {
    d1 = 0; d2 = 2;
    xzr = QIN(zr, max_n1); xzi = QIN(zr, max_n1 + 1);
    for (int k = 0; k < 2; k++) { vyr[k] = QIN(zr + 2 * (k - 1), max_n2); vyi[k] = QIN(zr + 2 * (k - 1), max_n2 + 1); }
  } else if (max_tr == 2) {
    d1 = 0; d2 = 1;
    xzr = QIN(zr, max_n1); xzi = QIN(zr, max_n1 + 1);
    for (int k = 0; k < 2; k++) { vyr[k] = QIN(zr + (k - 1), max_n2); vyi[k] = QIN(zr + (k - 1), max_n2 + 1); }
  } else {
    d1 = 2; d2 = 0;
    mid = max_tr;
    max_tr = 4 - max_tr;
    xzr = QIN(zr, max_n2); xzi = QIN(zr, max_n2 + 1);
    for (int k = 0; k < 2; k++) { vyr[k] = QIN(zr + 2 * (k - 1), max_n1); vyi[k] = QIN(zr + 2 * (k - 1), max_n1 + 1); }
  }
  {
    double base = 1e-17;
    base = base + xzr * xzr;
    base = base + xzi * xzi;
    float t = (float)sqrt(sqrt(base));
    float m = t * (float)(sqrt(t));
    m = 1 / m;
    xzr *= m; xzi *= m;
    for (int k = 0; k < 2; k++) {
      base = 1e-17;
      base = base + vyr[k] * vyr[k];
      base = base + vyi[k] * vyi[k];
      t = (float)sqrt(sqrt(base));
      m = t * (float)(sqrt(t));
      m = 1 / m;
      vyr[k] *= m; vyi[k] *= m;
    }
  }
  cpow_n(&xzr, &xzi, mid);
  for (int k = 0; k < 2; k++) cpow_n(&vyr[k], &vyi[k], max_tr);
  for (int k = 0; k < 2; k++) {
    vor[k] = vyr[k] * xzr - vyi[k] * xzi;
    voi[k] = vyr[k] * xzi + vyi[k] * xzr;
  }
  {
    float c, s;
    if (d2 == 1) {
      c = h->rom[XO_HROM_XP41 + (pidx << 1)];
      s = h->rom[XO_HROM_XP41 + (pidx << 1) + 1];
    } else {
      c = h->rom[XO_HROM_XP4 + (pidx << 1)];
      s = h->rom[XO_HROM_XP4 + (pidx << 1) + 1];
      if (d2 < d1) s = -s;
    }
    float tr_ = vor[0], ti_ = voi[0];
    vor[0] = (float)(c * tr_ - s * ti_);
    voi[0] = (float)(c * ti_ + s * tr_);
  }
  for (int k = 0; k < 2; k++) {
    h->qout[(col * 2 + k + 5) * 128 + 2 * band] += (float)(2.0f * vor[k]);
    h->qout[(col * 2 + k + 5) * 128 + 2 * band + 1] += (float)(2.0f * voi[k]);
  }
}

/* ixheaacd_hbe_post_anal_prod2 (:757-791) / _xprod2 (:1061-1248) */
static void prod2(hbe_t *h, int xprod, float p, const float *cs_theta) {
  norm_rows(h, h->xo[0], h->xo[1], 2);
  for (int b = h->xo[0]; b < h->xo[1]; b++) {
    int n1 = 0, n2 = 0;
    if (xprod) {
      double temp_fac = (2.0 * b + 1 - p) * 0.5;
      n1 = ((int)(temp_fac)) << 1;
      n2 = ((int)(temp_fac + p)) << 1;
    }
    for (int i = 0; i < 16; i++) {
      const float xzr = h->norm[(6 + i) * 128 + 2 * b], xzi = h->norm[(6 + i) * 128 + 2 * b + 1];
      for (int k = 0; k < 10; k++) {
        const float tr = h->norm[(1 + i + k) * 128 + 2 * b], ti = h->norm[(1 + i + k) * 128 + 2 * b + 1];
        h->qout[(1 + 2 * i + k) * 128 + 2 * b] += ((tr * xzr - ti * xzi) * 0.3333333f);
        h->qout[(1 + 2 * i + k) * 128 + 2 * b + 1] += ((tr * xzi + ti * xzr) * 0.3333333f);
      }
      if (!xprod) continue;
      const int zr = i + 6;
      float mag_zero = QIN(zr, 2 * b) * QIN(zr, 2 * b) + QIN(zr, 2 * b + 1) * QIN(zr, 2 * b + 1);
      float m1 = QIN(zr, n1) * QIN(zr, n1) + QIN(zr, n1 + 1) * QIN(zr, n1 + 1);
      float m2 = QIN(zr, n2) * QIN(zr, n2) + QIN(zr, n2 + 1) * QIN(zr, n2 + 1);
      float t = fmin_(m1, m2);
      float max_mag = 0;
      int max_n1 = 0, max_n2 = 0;
      if (t > 0) { max_mag = t; max_n1 = n1; max_n2 = n2; }
      if (!(max_mag > mag_zero && max_n1 >= 0 && max_n2 < 128)) continue;
      float zr_ = QIN(zr, max_n1), zi_ = QIN(zr, max_n1 + 1), vyr[2], vyi[2];
      float m = mag2(zr_, zi_);
      zr_ *= m; zi_ *= m;
      for (int k = 0; k < 2; k++) {
        float tr = QIN(zr - 1 + k, max_n2), ti = QIN(zr - 1 + k, max_n2 + 1);
        m = mag2(tr, ti);
        vyr[k] = tr * m;
        vyi[k] = ti * m;
      }
      float tr = vyr[0] * zr_ - vyi[0] * zi_, ti = vyr[0] * zi_ + vyi[0] * zr_;
      float tr1 = (float)(cs_theta[0] * tr - cs_theta[1] * ti);
      ti = (float)(cs_theta[0] * ti + cs_theta[1] * tr);
      h->qout[(i * 2 + 5) * 128 + 2 * b] += (float)(1.666666667f * tr1);
      h->qout[(i * 2 + 5) * 128 + 2 * b + 1] += (float)(1.666666667f * ti);
      tr = vyr[1] * zr_ - vyi[1] * zi_;
      ti = vyr[1] * zi_ + vyi[1] * zr_;
      h->qout[(i * 2 + 6) * 128 + 2 * b] += (float)(1.666666667f * tr);
      h->qout[(i * 2 + 6) * 128 + 2 * b + 1] += (float)(1.666666667f * ti);
    }
  }
}

/* ixheaacd_hbe_post_anal_prod3 (:793-1008) / _xprod3 (:1250-1473) */
static void prod3(hbe_t *h, int xprod, float p, int pidx) {
  const float *selc = h->rom + XO_HROM_SELCASE;
  for (int b = h->xo[1]; b < h->xo[2]; b++) {
    const int inp = (2 * b) / 3;
    const float *sel = selc + 8 * ((inp + 1) & 3), *sel1 = selc + 8 * (((inp + 1) & 3) + 1);
    const int rem = 2 * b - 3 * inp;
    for (int i = 0; i < 16; i++) {
      float vx[16], vc[16];
      float xzr, xzi, yr = 0, yi = 0;
      if (rem == 0 || rem == 1) {
        for (int m = 0; m < 4; m++) {
          const int r = i + 3 * m;
          float tr = QIN(r, 2 * inp), ti = QIN(r, 2 * inp + 1);
          float mg = mag3(tr, ti);
          vx[4 * m] = tr * mg;
          vx[4 * m + 1] = ti * mg;
          tr = QIN(r + 2, 2 * inp); ti = QIN(r + 2, 2 * inp + 1);
          float tr1 = sel[0] * tr + sel[1] * ti, ti1 = sel[2] * tr + sel[3] * ti;
          tr = QIN(r + 1, 2 * inp); ti = QIN(r + 1, 2 * inp + 1);
          tr1 += sel[4] * tr + sel[5] * ti;
          ti1 += sel[6] * tr + sel[7] * ti;
          tr1 *= 0.3984033437f;
          ti1 *= 0.3984033437f;
          mg = mag3(tr1, ti1);
          vx[4 * m + 2] = tr1 * mg;
          vx[4 * m + 3] = ti1 * mg;
        }
        const float tr = vx[8], ti = vx[9];
        xzr = tr * tr - ti * ti;
        xzi = tr * ti + ti * tr;
        for (int k = 0; k < 8; k++) {
          const float ar = vx[2 * k] * xzr - vx[2 * k + 1] * xzi, ai = vx[2 * k] * xzi + vx[2 * k + 1] * xzr;
          h->qout[(2 + 2 * i + k) * 128 + 2 * b] += (ar * 0.4714045f);
          h->qout[(2 + 2 * i + k) * 128 + 2 * b + 1] += (ai * 0.4714045f);
        }
      } else {
        for (int m = 0; m < 4; m++) {
          const int r = i + 3 * m;
          float tr1 = QIN(r, 2 * inp), ti1 = QIN(r, 2 * inp + 1);
          float tr = QIN(r, 2 * inp + 2), ti = QIN(r, 2 * inp + 3);
          float mg = mag3(tr, ti);
          vx[4 * m] = tr * mg;
          vx[4 * m + 1] = ti * mg;
          mg = mag3(tr1, ti1);
          vc[4 * m] = tr1 * mg;
          vc[4 * m + 1] = ti1 * mg;
          tr = QIN(r + 2, 2 * inp); ti = QIN(r + 2, 2 * inp + 1);
          tr1 = sel[0] * tr + sel[1] * ti;
          ti1 = sel[2] * tr + sel[3] * ti;
          tr = QIN(r + 1, 2 * inp); ti = QIN(r + 1, 2 * inp + 1);
          float cr = tr1 + sel[4] * tr + sel[5] * ti, ci = ti1 + sel[6] * tr + sel[7] * ti;
          tr = QIN(r + 2, 2 * inp + 2); ti = QIN(r + 2, 2 * inp + 3);
          tr1 = sel1[0] * tr + sel1[1] * ti;
          ti1 = sel1[2] * tr + sel1[3] * ti;
          tr = QIN(r + 1, 2 * inp + 2); ti = QIN(r + 1, 2 * inp + 3);
          float vr = tr1 + sel1[4] * tr + sel1[5] * ti, vi = ti1 + sel1[6] * tr + sel1[7] * ti;
          cr *= 0.3984033437f; ci *= 0.3984033437f;
          vr *= 0.3984033437f; vi *= 0.3984033437f;
          mg = mag3(vr, vi);
          vx[4 * m + 2] = vr * mg;
          vx[4 * m + 3] = vi * mg;
          mg = mag3(cr, ci);
          vc[4 * m + 2] = cr * mg;
          vc[4 * m + 3] = ci * mg;
        }
        float tr = vc[8], ti = vc[9], tr1 = vx[8], ti1 = vx[9];
        xzr = tr * tr - ti * ti;
        xzi = tr * ti + ti * tr;
        yr = tr1 * tr1 - ti1 * ti1;
        yi = tr1 * ti1 + ti1 * tr1;
        for (int k = 0; k < 8; k++) {
          float ar = vx[2 * k] * xzr - vx[2 * k + 1] * xzi, ai = vx[2 * k] * xzi + vx[2 * k + 1] * xzr;
          ar += vc[2 * k] * yr - vc[2 * k + 1] * yi;
          ai += vc[2 * k] * yi + vc[2 * k + 1] * yr;
          h->qout[(2 + 2 * i + k) * 128 + 2 * b] += (ar * 0.23570225f);
          h->qout[(2 + 2 * i + k) * 128 + 2 * b + 1] += (ai * 0.23570225f);
        }
      }
      if (xprod) xprod_proc_3(h, b, i, p, pidx);
    }
  }
}

/* ixheaacd_hbe_post_anal_prod4 (:1010-1059) / _xprod4 (:1475-1547) */
static void prod4(hbe_t *h, int xprod, float p, int pidx) {
  norm_rows(h, (h->xo[2] >> 1) - 1, h->xo[3], 4);
  for (int b = h->xo[2]; b < h->xo[3]; b++) {
    const int inp = b >> 1, ip = (b & 1) ? (inp + 1) : (inp - 1);
    for (int i = 0; i < 16; i++) {
      float xr = h->norm[(6 + i) * 128 + 2 * inp], xi = h->norm[(6 + i) * 128 + 2 * inp + 1];
      const float tr = xr, ti = xi;
      float t = xr * xr - xi * xi;
      xi = xr * xi + xi * xr;
      xr = tr * t - ti * xi;
      xi = tr * xi + ti * t;
      for (int k = 0; k < 6; k++) {
        const float a = h->norm[(i + 2 * k) * 128 + 2 * ip], bi = h->norm[(i + 2 * k) * 128 + 2 * ip + 1];
        const float orr = a * xr - bi * xi, oi = a * xi + bi * xr;
        h->qout[(3 + 2 * i + k) * 128 + 2 * b] += (orr * 0.6666667f);
        h->qout[(3 + 2 * i + k) * 128 + 2 * b + 1] += (oi * 0.6666667f);
      }
      if (xprod) xprod_proc_4(h, b, i, p, pidx);
    }
  }
}

int xo_esbr_hbe_apply(const float *rom, const int32_t *cfg, float *state, const float *qmf_re, const float *qmf_im,
                      float *pv_re, float *pv_im) {
  static hbe_t hs; /* single-threaded test infrastructure */
  hbe_t *h = &hs;
  memset(h, 0, sizeof(*h));
  h->rom = rom;
  h->S = cfg[XO_HBE_SYNTH_SIZE];
  h->k_start = cfg[XO_HBE_K_START];
  h->start_band = cfg[XO_HBE_START_BAND];
  h->end_band = cfg[XO_HBE_END_BAND];
  h->max_stretch = cfg[XO_HBE_MAX_STRETCH];
  for (int i = 0; i < 6; i++) h->xo[i] = cfg[XO_HBE_XOVER + i];
  const int S = h->S, pitch = cfg[XO_HBE_PITCH];
  if (cfg[XO_HBE_USF4] || !(S == 4 || S == 8 || S == 12 || S == 16 || S == 20)) return -2;
  if (h->k_start < 0) return -1; /* polyphase.c:187 */
  if (h->k_start + S > 32 || h->start_band < 0 || h->end_band > 64 || h->start_band > h->end_band) return -2;
  for (int i = 0; i < 4; i++)
    if (h->xo[i] < 0 || h->xo[i] > 64) return -2;
  /* hbe_trans.c:235-238: the first samples of the time buffer are the previous call's last ones */
  memcpy(h->inbuf, state + XO_HBE_ST_TAIL, S * sizeof(float));
  if (S != 20) { /* synth_size 20: the reference re-initialises (clears both histories) on every call, hbe_trans.c:240-248 */
    memcpy(h->synth_buf, state + XO_HBE_ST_SYNTH, 18 * S * sizeof(float));
    memcpy(h->analy_buf, state + XO_HBE_ST_ANAL, 18 * S * sizeof(float));
  }
  real_synth_filt(h, qmf_re, qmf_im);
  memcpy(h->qin, state + XO_HBE_ST_QIN, 12 * 128 * sizeof(float)); /* :254-258 */
  complex_anal_filt(h);
  memcpy(h->qout, state + XO_HBE_ST_QOUT, 10 * 128 * sizeof(float)); /* :263-273: shift by no_bins, clear the rest */
  {
    const float p = (float)(pitch * 0.08333333333333); /* :1557-1558, 2:1 system */
    if (p < 1.0f) {
      if (2 <= h->max_stretch) prod2(h, 0, p, NULL);
      if (3 <= h->max_stretch) prod3(h, 0, p, 0);
      if (4 <= h->max_stretch) {
        if (h->xo[2] <= 1) return (int)0x80000000;
        prod4(h, 0, p, 0);
      }
    } else {
      if (2 <= h->max_stretch) prod2(h, 1, p, rom + XO_HROM_XP2 + (pitch << 1));
      if (3 <= h->max_stretch) prod3(h, 1, p, pitch);
      if (4 <= h->max_stretch) {
        if (h->xo[2] <= 1) return (int)0x80000000;
        prod4(h, 1, p, pitch);
      }
    }
  }
  const float *pc = rom + XO_HROM_PVCOS, *ps = rom + XO_HROM_PVSIN;
  for (int i = 0; i < 32; i++)
    for (int b = h->start_band; b < h->end_band; b++) { /* :280-294 */
      const float a = h->qout[i * 128 + 2 * b], c = h->qout[i * 128 + 2 * b + 1];
      pv_re[i * 64 + b] = (float)(a * pc[b] - c * ps[b]);
      pv_im[i * 64 + b] = (float)(a * ps[b] + c * pc[b]);
    }
  memcpy(state + XO_HBE_ST_TAIL, h->inbuf + 32 * S, S * sizeof(float));
  memcpy(state + XO_HBE_ST_SYNTH, h->synth_buf, 18 * S * sizeof(float));
  memcpy(state + XO_HBE_ST_ANAL, h->analy_buf, 18 * S * sizeof(float));
  memcpy(state + XO_HBE_ST_QIN, h->qin + 16 * 128, 12 * 128 * sizeof(float));
  memcpy(state + XO_HBE_ST_QOUT, h->qout + 32 * 128, 10 * 128 * sizeof(float));
  return 0;
}

void xo_esbr_hbe_apply_batch(const float *rom, const int32_t *cfg, float *state, const float *qmf_re, const float *qmf_im,
                             float *pv_re, float *pv_im, int32_t *err, int n) {
  for (int u = 0; u < n; u++) {
    int e = xo_esbr_hbe_apply(rom, cfg + (size_t)u * XO_HBE_CFG_WORDS, state + (size_t)u * XO_HBE_ST_WORDS,
                              qmf_re + (size_t)u * 2048, qmf_im + (size_t)u * 2048, pv_re + (size_t)u * 2048,
                              pv_im + (size_t)u * 2048);
    if (err) err[u] = e;
  }
}
