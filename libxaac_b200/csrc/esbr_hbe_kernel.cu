// esbr_hbe_kernel.cu — the QMF-domain harmonic transposer of the float eSBR decoder for sm_100a (B200).
//
// One warp owns one unit (one channel of one frame).  Replaces
//   ixheaacd_qmf_hbe_apply            decoder/ixheaacd_hbe_trans.c:224-296
//   ixheaacd_real_synth_filt          decoder/ixheaacd_esbr_polyphase.c:157-274   (critically sampled real synthesis bank)
//   ixheaacd_complex_anal_filt        decoder/ixheaacd_esbr_polyphase.c:48-155    (2x complex analysis bank)
//   ixheaacd_hbe_post_anal_process    decoder/ixheaacd_hbe_trans.c:1549-1606 with prod2/3/4, xprod2/3/4, xprod_proc_3/4,
//                                     norm_qmf_in_buf_2/4 (:298-1547)
//   ixheaac_real_synth_fft_p2/_p3, ixheaac_cmplx_anal_fft_p2/_p3   common/ixheaac_esbr_fft.c
// for the 2:1 system (32 QMF columns per call).  This file is compiled with -fmad=false: the reference build has no FMA
// and evaluates float expressions in float and double ones in double, so every operation below is one IEEE rounding in
// the reference's order and the float output is bit-identical (the only libm call, cbrt, is correctly rounded to float
// on both sides in all but ~2^-28 of the cases).
//
// Phases and lane maps:
//   A  real synthesis bank     lane = QMF column (32 independent modulations; the window is a 10-tap FIR across columns)
//   B  complex analysis bank   window: lane = sample; direct-form modulation (bank size 40): lane = (column, half of the bins),
//                              two bins per pass sharing the sample loads; FFT banks: lane = column
//   C  stretch-2/3/4 products  lane = (band parity, column).  The reference accumulates the contributions of one output
//                              cell in column order; with every lane walking its own column's taps in DESCENDING tap
//                              order in lock step, the adds into a cell happen in ascending column order — the same sums.
//   D  rotation + store        lane = (row, band) over 8-band chunks
// The time signal, the bank histories, the 28-row analysis matrix and the normalised matrices stay in shared memory.
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "kernels.h"

namespace xb {

#define XB_DEV __device__ __forceinline__
#define XB_NOINL __device__ __noinline__

constexpr int kHbWarps = 8;
// CTA-shared tables (float words)
constexpr int kHtWin = 0;         // ixheaac_sub_samp_qmf_window_coeff[1560]
constexpr int kHtSyn20 = 1560;    // ixheaac_synth_cos_table_kl_20, rows 0..30 x 20
constexpr int kHtCosTrans = 2360; // ixheaac_cos_table_trans_qmf[7][64]
constexpr int kHtAna40 = 2808;    // ixheaac_analy_cos_sin_table_kl_40 as 40 rows of stride 82 (two rows 20 apart: distinct banks)
constexpr int kHtWords = 2808 + 40 * 82;
// per-warp regions (float words)
constexpr int kHbE = 1024;   // time signal (<= 51 S - 1 words); phase C: NORM4 [26][37] (<= 18 bands)
constexpr int kHbR1 = 2688;  // A: V [41][<=41] + LOC [20][33] | B: U [80][17] | C: NORM2 [26][73] + OUTC [42][17]
constexpr int kHbQW = 82;    // row stride of the analysis matrix (<= 80 words used)
constexpr int kHbQ = 28 * kHbQW;
constexpr int kHbWarpWords = kHbE + kHbR1 + kHbQ;
constexpr int kNW = 73;      // row stride of the stretch-2 normalised matrix (<= 36 bands; odd: lane = column walks rows)
constexpr int kN4W = 37;     // row stride of the stretch-4 normalised matrix (<= 18 bands)
constexpr int kNRows = 26;   // rows 0..25 of the normalised matrices are the ones the products read
constexpr int kOW = 17;      // row stride of the 8-band output chunk (16 words used)

struct cf { float r, i; };

// ---- radix-4 FFT of common/ixheaac_esbr_fft.c on a lane-private array (FFT banks only; bank size 40 is direct form) ----
XB_DEV void rot_a(cf &x, float wc, float ws) { float t = x.r * wc + x.i * ws; x.i = -(x.r * ws) + x.i * wc; x.r = t; }
XB_DEV void rot_b(cf &x, float w3, float w6) { float t = x.r * w6 - x.i * w3; x.i = x.r * w3 + x.i * w6; x.r = t; }
XB_DEV void rot_c(cf &x, float w3, float w6) { float t = -(x.r * w3) - x.i * w6; x.i = -(x.r * w6) + x.i * w3; x.r = t; }
XB_DEV void bfly4(cf x0, cf x1, cf x2, cf x3, bool alt, float *d, int st) {
  x0.r = x0.r + x2.r; x0.i = x0.i + x2.i;
  x2.r = x0.r - (x2.r * 2); x2.i = x0.i - (x2.i * 2);
  x1.r = x1.r + x3.r;
  if (!alt) { x1.i = x1.i + x3.i; x3.r = x1.r - (x3.r * 2); x3.i = x1.i - (x3.i * 2); }
  else { x1.i = x1.i - x3.i; x3.r = x1.r - (x3.r * 2); x3.i = x1.i + (x3.i * 2); }
  x0.r = x0.r + x1.r; x0.i = x0.i + x1.i;
  x1.r = x0.r - (x1.r * 2); x1.i = x0.i - (x1.i * 2);
  x2.r = x2.r - x3.i; x2.i = x2.i + x3.r;
  x3.i = x2.r + (x3.i * 2); x3.r = x2.i - (x3.r * 2);
  d[0] = x0.r; d[1] = x0.i; d[st] = x2.r; d[st + 1] = x2.i;
  d[2 * st] = x1.r; d[2 * st + 1] = x1.i; d[3 * st] = x3.i; d[3 * st + 1] = x3.r;
}
XB_DEV unsigned dig_rev(unsigned v, int m) {
  v = ((v & 0x33333333u) << 2) | ((v & ~0x33333333u) >> 2);
  v = ((v & 0x0F0F0F0Fu) << 4) | ((v & ~0x0F0F0F0Fu) >> 4);
  v = ((v & 0x00FF00FFu) << 8) | ((v & ~0x00FF00FFu) >> 8);
  return v >> m;
}
XB_DEV int ilog2(int n) { return 31 - __clz(n); }

XB_NOINL void fft_rest(const float *tw, float *y, int npoints) {  // esbr_fft.c:95-535
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
        bool alt = false;
        if (jj > 0) {
          rot_a(x1, __ldg(tw + j), __ldg(tw + j + 257));
          if (j <= sec) {
            rot_a(x2, __ldg(tw + 2 * j), __ldg(tw + 2 * j + 257));
            rot_a(x3, __ldg(tw + 3 * j), __ldg(tw + 3 * j + 257));
          } else if (j <= (nd >> 1)) {
            rot_a(x2, __ldg(tw + 2 * j), __ldg(tw + 2 * j + 257));
            rot_b(x3, __ldg(tw + 3 * j - 256), __ldg(tw + 3 * j + 1));
          } else if (j <= sec * 2) {
            rot_b(x2, __ldg(tw + 2 * j - 256), __ldg(tw + 2 * j + 1));
            rot_b(x3, __ldg(tw + 3 * j - 256), __ldg(tw + 3 * j + 1));
          } else {
            rot_b(x2, __ldg(tw + 2 * j - 256), __ldg(tw + 2 * j + 1));
            rot_c(x3, __ldg(tw + 3 * j - 512), __ldg(tw + 3 * j - 512 + 257));
            alt = true;
          }
        }
        bfly4(x0, x1, x2, x3, alt, d, 2 * del);
      }
    }
    nodespacing >>= 2;
    del <<= 2;
    in_loop_cnt >>= 2;
  }
  if (not_power_4) {
    nodespacing <<= 1;
    for (int t = 0; t < del; t++) {
      const int q = t < del / 2 ? t : t - del / 2;
      const float w1 = __ldg(tw + q * nodespacing), w4 = __ldg(tw + q * nodespacing + 257);
      float *p = y + 2 * t;
      cf x0 = {p[0], p[1]}, x1 = {p[2 * del], p[2 * del + 1]};
      if (t < del / 2) rot_a(x1, w1, w4);
      else rot_b(x1, w1, w4);
      p[2 * del] = x0.r - x1.r;
      p[2 * del + 1] = x0.i - x1.i;
      p[0] = x0.r + x1.r;
      p[1] = x0.i + x1.i;
    }
  }
}
XB_NOINL void real_synth_fft_p2(const float *tw, const float *x, float *y, int npoints) {  // :42
  const int lg = ilog2(npoints), not_power_4 = lg & 1;
  const int shift = (31 - 1 - lg) + 1 - 16;
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
XB_NOINL void cmplx_anal_fft_p2(const float *tw, const float *x, float *y, int npoints) {  // :537
  const int lg = ilog2(npoints), not_power_4 = lg & 1;
  const int shift = (31 - 1 - lg) + 1 - 16;
  for (int i = 0; i < npoints; i += 4) {
    int h2 = (int)dig_rev((unsigned)i, shift);
    if (not_power_4) h2 = (h2 + 1) & ~1;
    const float *inp = x + h2;
    const int st = npoints >> 1;
    cf x0 = {inp[0], inp[1]}, x1 = {inp[st], inp[st + 1]}, x2 = {inp[2 * st], inp[2 * st + 1]},
       x3 = {inp[3 * st], inp[3 * st + 1]};
    bfly4(x0, x1, x2, x3, false, y + 2 * i, 2);
  }
  fft_rest(tw, y, npoints);
}
XB_DEV void fft3(const float *inp, float *op) {  // :1048
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
XB_DEV void tw3(float *x, const float *wr, int n) {  // :1110 / :1171
  x += 2;
  for (int i = 0; i < n; i++) {
    for (int q = 0; q < 2; q++) {
      float w0 = __ldg(wr), w1 = __ldg(wr + 1);
      float t = x[0] * w0 + x[1] * w1;
      x[1] = -x[0] * w1 + x[1] * w0;
      x[0] = t;
      wr += 2;
      x += 2;
    }
    x += 2;
  }
}
XB_NOINL void real_synth_fft_p3(const float *rom, const float *x_in, float *x_out) {  // :1084, 24 points
  float x_3[8], y_3[16], y[48], x[48];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 8; j++) x_3[j] = x_in[3 * j + i];
    real_synth_fft_p2(rom + kHromFftTw, x_3, y_3, 8);
    for (int j = 0; j < 16; j += 2) {
      x[3 * j + 2 * i] = y_3[j];
      x[3 * j + 2 * i + 1] = y_3[j + 1];
    }
  }
  tw3(x, rom + kHromTw24, 8);
  for (int i = 0; i < 8; i++) fft3(x + 6 * i, y + 6 * i);
  const float *py = y;
  for (int i = 0; i < 16; i += 2) {
    x_out[i] = *py++; x_out[i + 1] = *py++;
    x_out[16 + i] = *py++; x_out[16 + i + 1] = *py++;
    x_out[32 + i] = *py++; x_out[32 + i + 1] = *py++;
  }
}
XB_NOINL void cmplx_anal_fft_p3(const float *rom, float *x_in, float *x_out) {  // :1148, 48 points, in place on x_in
  float x_3[32], y_3[32], y[96];
  for (int i = 0; i < 6; i += 2) {
    for (int j = 0; j < 32; j += 2) {
      x_3[j] = x_in[3 * j + i];
      x_3[j + 1] = x_in[3 * j + i + 1];
    }
    cmplx_anal_fft_p2(rom + kHromFftTw, x_3, y_3, 16);
    for (int j = 0; j < 32; j += 2) {
      x_in[3 * j + i] = y_3[j];
      x_in[3 * j + i + 1] = y_3[j + 1];
    }
  }
  tw3(x_in, rom + kHromTw48, 16);
  for (int i = 0; i < 16; i++) fft3(x_in + 6 * i, y + 6 * i);
  const float *py = y;
  for (int i = 0; i < 32; i += 2) {
    x_out[i] = *py++; x_out[i + 1] = *py++;
    x_out[32 + i] = *py++; x_out[32 + i + 1] = *py++;
    x_out[64 + i] = *py++; x_out[64 + i + 1] = *py++;
  }
}

XB_DEV int win_off(int len) {  // ixheaacd_map_prot_filter, hbe_trans.c:70
  return len == 4 ? 0 : len == 8 ? 40 : len == 12 ? 120 : len == 16 ? 240 : len == 20 ? 400 : len == 24 ? 600
       : len == 32 ? 840 : len == 40 ? 1160 : 0;
}
XB_DEV int syncos_off(int s) { return s == 4 ? 0 : s == 8 ? 16 : s == 12 ? 48 : 96; }
XB_DEV int anacs_off(int s) { return s == 4 ? 0 : s == 8 ? 32 : s == 12 ? 96 : 192; }

// ---- magnitude normalisations (hbe_trans.c:312-322, 351-358, 832-835) ----
XB_DEV float mag2(float xr, float xi) {
  double base = 1e-17;
  float t = xr * xr;
  base = base + (double)t;
  float t2 = xi * xi;
  base = base + (double)t2;
  float m = (float)(1.0 / base);
  return (float)sqrt(sqrt((double)m));
}
XB_DEV float mag3(float xr, float xi) {
  float t = xr * xr, t2 = xi * xi;
  double b1 = 1e-17 + (double)t;
  b1 = b1 + (double)t2;
  float q = 1.0f / (float)b1;
  return (float)cbrt((double)q);
}
XB_DEV float mag4(float xr, float xi) {
  double base = 1e-17;
  float t = xr * xr;
  base = base + (double)t;
  t = xi * xi;
  base = base + (double)t;
  t = (float)sqrt(sqrt(base));
  float m = t * sqrtf(t);
  return 1.0f / m;
}
XB_DEV void cpow_n(float &xr, float &xi, int n) {  // :521-527
  const float tr = xr, ti = xi;
  for (int q = 0; q < n - 1; q++) {
    float tmp = xr;
    xr = xr * tr - xi * ti;
    xi = tmp * ti + xi * tr;
  }
}

// the 28-row analysis matrix: row r holds the reference's qmf_in_buf[r][lo .. lo + 4 S); everything else is zero there.
// `fl` may run past a row end like the reference's own pointer arithmetic in the cross-product search.
struct QinView {
  const float *q;
  int lo, n;
  XB_DEV float at(int row, int fl) const {
    const int idx = row * 128 + fl;
    const int r = idx >> 7, c = (idx & 127) - lo;
    return (r >= 0 && r < 28 && c >= 0 && c < n) ? q[r * kHbQW + c] : 0.0f;
  }
};

struct XAdd { float r0, i0, r1, i1; bool on; };  // cross-product contributions to rows 2 col + 5 and 2 col + 6

// hbe_trans.c:1105-1243 (inside ixheaacd_hbe_post_anal_xprod2).  N2 = the stretch-2 normalised matrix (bands nlo..nhi, rows
// 1..25): a normalised cell the search needs is the same expression the matrix already holds (x * mag2(x)), so it is read back
// instead of being recomputed when it lies inside the matrix.
XB_DEV XAdd xprod2(const QinView &Q, const float *N2, int nlo, int nhi, int b, int col, float p, const float *cs_theta) {
  XAdd a = {0, 0, 0, 0, false};
  double temp_fac = (2.0 * b + 1 - (double)p) * 0.5;
  const int n1 = ((int)(temp_fac)) << 1, n2 = ((int)(temp_fac + (double)p)) << 1;
  const int zr = col + 6;
  const float z0 = Q.at(zr, 2 * b), z1 = Q.at(zr, 2 * b + 1);
  const float a0 = Q.at(zr, n1), a1 = Q.at(zr, n1 + 1), b0 = Q.at(zr, n2), b1 = Q.at(zr, n2 + 1);
  float mag_zero = z0 * z0 + z1 * z1;
  float m1 = a0 * a0 + a1 * a1;
  float m2 = b0 * b0 + b1 * b1;
  float t = m1 < m2 ? m1 : m2;
  float max_mag = 0;
  int max_n1 = 0, max_n2 = 0;
  if (t > 0) { max_mag = t; max_n1 = n1; max_n2 = n2; }
  if (!(max_mag > mag_zero && max_n1 >= 0 && max_n2 < 128)) return a;
  float zr_, zi_, vyr[2], vyi[2];
  const int ba = max_n1 >> 1, bb = max_n2 >> 1;
  if (ba >= nlo && ba <= nhi) {
    zr_ = N2[zr * kNW + 2 * (ba - nlo)];
    zi_ = N2[zr * kNW + 2 * (ba - nlo) + 1];
  } else {
    const float m = mag2(a0, a1);
    zr_ = a0 * m; zi_ = a1 * m;
  }
#pragma unroll
  for (int k = 0; k < 2; k++) {
    if (bb >= nlo && bb <= nhi) {
      vyr[k] = N2[(zr - 1 + k) * kNW + 2 * (bb - nlo)];
      vyi[k] = N2[(zr - 1 + k) * kNW + 2 * (bb - nlo) + 1];
    } else {
      float tr = Q.at(zr - 1 + k, max_n2), ti = Q.at(zr - 1 + k, max_n2 + 1);
      const float m = mag2(tr, ti);
      vyr[k] = tr * m;
      vyi[k] = ti * m;
    }
  }
  float tr = vyr[0] * zr_ - vyi[0] * zi_, ti = vyr[0] * zi_ + vyi[0] * zr_;
  const float c0 = __ldg(cs_theta), c1 = __ldg(cs_theta + 1);
  float tr1 = c0 * tr - c1 * ti;
  ti = c0 * ti + c1 * tr;
  a.r0 = 1.666666667f * tr1;
  a.i0 = 1.666666667f * ti;
  tr = vyr[1] * zr_ - vyi[1] * zi_;
  ti = vyr[1] * zi_ + vyi[1] * zr_;
  a.r1 = 1.666666667f * tr;
  a.i1 = 1.666666667f * ti;
  a.on = true;
  return a;
}

// ixheaacd_hbe_xprod_proc_3 (hbe_trans.c:371-553)
XB_NOINL XAdd xprod3(const float *q, int lo, int n, const float *rom, int band, int col, float p, int pidx) {
  QinView Q = {q, lo, n};
  XAdd a = {0, 0, 0, 0, false};
  const int inp_band = 2 * band / 3, zr = col + 6;
  float mag_zero = Q.at(zr, 2 * inp_band) * Q.at(zr, 2 * inp_band) + Q.at(zr, 2 * inp_band + 1) * Q.at(zr, 2 * inp_band + 1);
  float max_mag = 0;
  int max_n1 = 0, max_n2 = 0, max_tr = 0;
  for (int tr = 1; tr < 3; tr++) {
    float f = 2.0f * band + 1 - tr * p;
    double temp_fac = (double)f * 0.3333334;
    int n1 = (int)(temp_fac), n2 = (int)(temp_fac + (double)p);
    float m1 = Q.at(zr, 2 * n1) * Q.at(zr, 2 * n1) + Q.at(zr, 2 * n1 + 1) * Q.at(zr, 2 * n1 + 1);
    float m2 = Q.at(zr, 2 * n2) * Q.at(zr, 2 * n2) + Q.at(zr, 2 * n2 + 1) * Q.at(zr, 2 * n2 + 1);
    float t = m1 < m2 ? m1 : m2;
    if (t > max_mag) { max_mag = t; max_tr = tr; max_n1 = n1; max_n2 = n2; }
  }
  if (!(max_mag > mag_zero && max_n1 >= 0 && max_n2 < 64)) return a;
  float vyr[2], vyi[2], vor[2], voi[2], d1, d2, xzr, xzi;
  int mid = 3 - max_tr, na, nb;
  if (max_tr == 1) { d1 = 0; d2 = 1.5f; na = max_n1; nb = max_n2; }
  else { d1 = 1.5f; d2 = 0; mid = max_tr; max_tr = 3 - max_tr; na = max_n2; nb = max_n1; }
  xzr = Q.at(zr, 2 * na);
  xzi = Q.at(zr, 2 * na + 1);
  {
    const int idx = ((nb & 3) + 1) & 3;
    const float cr0 = __ldg(rom + kHromInterp + 2 * idx), ci0 = __ldg(rom + kHromInterp + 2 * idx + 1), cr1 = cr0, ci1 = -ci0;
    vyr[1] = Q.at(zr, 2 * nb);
    vyi[1] = Q.at(zr, 2 * nb + 1);
    float tr_ = Q.at(zr - 2, 2 * nb), ti_ = Q.at(zr - 2, 2 * nb + 1);
    vyr[0] = cr1 * tr_ - ci1 * ti_;
    vyi[0] = ci1 * tr_ + cr1 * ti_;
    tr_ = Q.at(zr - 1, 2 * nb); ti_ = Q.at(zr - 1, 2 * nb + 1);
    vyr[0] += cr0 * tr_ - ci0 * ti_;
    vyi[0] += ci0 * tr_ + cr0 * ti_;
  }
  float m = mag3(xzr, xzi);
  xzr *= m; xzi *= m;
  for (int k = 0; k < 2; k++) {
    m = mag3(vyr[k], vyi[k]);
    vyr[k] *= m; vyi[k] *= m;
  }
  cpow_n(xzr, xzi, mid);
  for (int k = 0; k < 2; k++) cpow_n(vyr[k], vyi[k], max_tr);
  for (int k = 0; k < 2; k++) {
    vor[k] = vyr[k] * xzr - vyi[k] * xzi;
    voi[k] = vyr[k] * xzi + vyi[k] * xzr;
  }
  {
    float c = __ldg(rom + kHromXp3 + (pidx << 1)), s = __ldg(rom + kHromXp3 + (pidx << 1) + 1);
    if (d2 < d1) s = -s;
    float tr_ = vor[0], ti_ = voi[0];
    vor[0] = c * tr_ - s * ti_;
    voi[0] = c * ti_ + s * tr_;
  }
  a.r0 = 1.8856f * vor[0]; a.i0 = 1.8856f * voi[0];
  a.r1 = 1.8856f * vor[1]; a.i1 = 1.8856f * voi[1];
  a.on = true;
  return a;
}

// ixheaacd_hbe_xprod_proc_4 (hbe_trans.c:555-755)
XB_NOINL XAdd xprod4(const float *q, int lo, int n, const float *rom, int band, int col, float p, int pidx) {
  QinView Q = {q, lo, n};
  XAdd a = {0, 0, 0, 0, false};
  const int inp_band = band >> 1, zr = col + 6;
  float mag_zero = Q.at(zr, 2 * inp_band) * Q.at(zr, 2 * inp_band) + Q.at(zr, 2 * inp_band + 1) * Q.at(zr, 2 * inp_band + 1);
  float max_mag = 0;
  int max_n1 = 0, max_n2 = 0, max_tr = 0;
  for (int tr = 1; tr < 4; tr++) {
    float tp = tr * p;
    double temp_fac = (2.0 * band + 1 - (double)tp) * 0.25;
    int n1 = ((int)(temp_fac)) << 1, n2 = ((int)(temp_fac + (double)p)) << 1;
    float m1 = Q.at(zr, n1) * Q.at(zr, n1) + Q.at(zr, n1 + 1) * Q.at(zr, n1 + 1);
    float m2 = Q.at(zr, n2) * Q.at(zr, n2) + Q.at(zr, n2 + 1) * Q.at(zr, n2 + 1);
    float t = m1 < m2 ? m1 : m2;
    if (t > max_mag) { max_mag = t; max_tr = tr; max_n1 = n1; max_n2 = n2; }
  }
  if (!(max_mag > mag_zero && max_n1 >= 0 && max_n2 < 128)) return a;
  float vyr[2], vyi[2], vor[2], voi[2], d1, d2, xzr, xzi;
  int mid = 4 - max_tr;
  if (max_tr == 1) {
    d1 = 0; d2 = 2;
    xzr = Q.at(zr, max_n1); xzi = Q.at(zr, max_n1 + 1);
    for (int k = 0; k < 2; k++) { vyr[k] = Q.at(zr + 2 * (k - 1), max_n2); vyi[k] = Q.at(zr + 2 * (k - 1), max_n2 + 1); }
  } else if (max_tr == 2) {
    d1 = 0; d2 = 1;
    xzr = Q.at(zr, max_n1); xzi = Q.at(zr, max_n1 + 1);
    for (int k = 0; k < 2; k++) { vyr[k] = Q.at(zr + (k - 1), max_n2); vyi[k] = Q.at(zr + (k - 1), max_n2 + 1); }
  } else {
    d1 = 2; d2 = 0;
    mid = max_tr;
    max_tr = 4 - max_tr;
    xzr = Q.at(zr, max_n2); xzi = Q.at(zr, max_n2 + 1);
    for (int k = 0; k < 2; k++) { vyr[k] = Q.at(zr + 2 * (k - 1), max_n1); vyi[k] = Q.at(zr + 2 * (k - 1), max_n1 + 1); }
  }
  float m = mag4(xzr, xzi);
  xzr *= m; xzi *= m;
  for (int k = 0; k < 2; k++) {
    m = mag4(vyr[k], vyi[k]);
    vyr[k] *= m; vyi[k] *= m;
  }
  cpow_n(xzr, xzi, mid);
  for (int k = 0; k < 2; k++) cpow_n(vyr[k], vyi[k], max_tr);
  for (int k = 0; k < 2; k++) {
    vor[k] = vyr[k] * xzr - vyi[k] * xzi;
    voi[k] = vyr[k] * xzi + vyi[k] * xzr;
  }
  {
    float c, s;
    if (d2 == 1) {
      c = __ldg(rom + kHromXp41 + (pidx << 1));
      s = __ldg(rom + kHromXp41 + (pidx << 1) + 1);
    } else {
      c = __ldg(rom + kHromXp4 + (pidx << 1));
      s = __ldg(rom + kHromXp4 + (pidx << 1) + 1);
      if (d2 < d1) s = -s;
    }
    float tr_ = vor[0], ti_ = voi[0];
    vor[0] = c * tr_ - s * ti_;
    voi[0] = c * ti_ + s * tr_;
  }
  a.r0 = 2.0f * vor[0]; a.i0 = 2.0f * voi[0];
  a.r1 = 2.0f * vor[1]; a.i1 = 2.0f * voi[1];
  a.on = true;
  return a;
}

// stretch-3 column vectors (hbe_trans.c:813-866 / :898-978): vx (and vc for the two-band case), 8 complex values each
XB_NOINL void prod3_vectors(const float *q, int lo, int n, const float *rom, int b, int i, float *vx, float *vc, bool &two) {
  QinView Q = {q, lo, n};
  const int inp = (2 * b) / 3, rem = 2 * b - 3 * inp;
  float sel[8], sel1[8];
#pragma unroll
  for (int t = 0; t < 8; t++) {
    sel[t] = __ldg(rom + kHromSelCase + 8 * ((inp + 1) & 3) + t);
    sel1[t] = __ldg(rom + kHromSelCase + 8 * (((inp + 1) & 3) + 1) + t);
  }
  two = !(rem == 0 || rem == 1);
  if (!two) {
    for (int m = 0; m < 4; m++) {
      const int r = i + 3 * m;
      float tr = Q.at(r, 2 * inp), ti = Q.at(r, 2 * inp + 1);
      float mg = mag3(tr, ti);
      vx[4 * m] = tr * mg;
      vx[4 * m + 1] = ti * mg;
      tr = Q.at(r + 2, 2 * inp); ti = Q.at(r + 2, 2 * inp + 1);
      float tr1 = sel[0] * tr + sel[1] * ti, ti1 = sel[2] * tr + sel[3] * ti;
      tr = Q.at(r + 1, 2 * inp); ti = Q.at(r + 1, 2 * inp + 1);
      tr1 += sel[4] * tr + sel[5] * ti;
      ti1 += sel[6] * tr + sel[7] * ti;
      tr1 *= 0.3984033437f;
      ti1 *= 0.3984033437f;
      mg = mag3(tr1, ti1);
      vx[4 * m + 2] = tr1 * mg;
      vx[4 * m + 3] = ti1 * mg;
    }
  } else {
    for (int m = 0; m < 4; m++) {
      const int r = i + 3 * m;
      float tr1 = Q.at(r, 2 * inp), ti1 = Q.at(r, 2 * inp + 1);
      float tr = Q.at(r, 2 * inp + 2), ti = Q.at(r, 2 * inp + 3);
      float mg = mag3(tr, ti);
      vx[4 * m] = tr * mg;
      vx[4 * m + 1] = ti * mg;
      mg = mag3(tr1, ti1);
      vc[4 * m] = tr1 * mg;
      vc[4 * m + 1] = ti1 * mg;
      tr = Q.at(r + 2, 2 * inp); ti = Q.at(r + 2, 2 * inp + 1);
      tr1 = sel[0] * tr + sel[1] * ti;
      ti1 = sel[2] * tr + sel[3] * ti;
      tr = Q.at(r + 1, 2 * inp); ti = Q.at(r + 1, 2 * inp + 1);
      float cr = tr1 + sel[4] * tr + sel[5] * ti, ci = ti1 + sel[6] * tr + sel[7] * ti;
      tr = Q.at(r + 2, 2 * inp + 2); ti = Q.at(r + 2, 2 * inp + 3);
      tr1 = sel1[0] * tr + sel1[1] * ti;
      ti1 = sel1[2] * tr + sel1[3] * ti;
      tr = Q.at(r + 1, 2 * inp + 2); ti = Q.at(r + 1, 2 * inp + 3);
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
  }
}

__global__ void __launch_bounds__(kHbWarps * 32) esbr_hbe_kernel(EsbrHbeArgs p) {
  extern __shared__ __align__(16) float hb_smem[];
  float *tabs = hb_smem;
  {
    const float *rom = p.rom;
    for (int i = threadIdx.x; i < 1560; i += blockDim.x) tabs[kHtWin + i] = __ldg(rom + kHromWin + i);
    for (int i = threadIdx.x; i < 800; i += blockDim.x) tabs[kHtSyn20 + i] = __ldg(rom + kHromSyn20 + i);
    for (int i = threadIdx.x; i < 448; i += blockDim.x) tabs[kHtCosTrans + i] = __ldg(rom + kHromCosTrans + i);
    for (int i = threadIdx.x; i < 3200; i += blockDim.x) tabs[kHtAna40 + (i / 80) * 82 + (i % 80)] = __ldg(rom + kHromAna40 + i);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *W = hb_smem + kHtWords + warp * kHbWarpWords;
  float *E = W, *R1 = W + kHbE, *QM = W + kHbE + kHbR1;
  const long long warps_total = (long long)gridDim.x * kHbWarps;
  const float *rom = p.rom;
  for (long long u = (long long)blockIdx.x * kHbWarps + warp; u < p.n_units; u += warps_total) {
    __syncwarp();
    int cw = lane < 16 ? __ldg(p.cfg + u * kHbeCfgWords + lane) : 0;
    const int S = __shfl_sync(0xffffffffu, cw, kHbeSynthSize), ks = __shfl_sync(0xffffffffu, cw, kHbeKStart);
    const int sb = __shfl_sync(0xffffffffu, cw, kHbeStartBand), eb = __shfl_sync(0xffffffffu, cw, kHbeEndBand);
    const int ms = __shfl_sync(0xffffffffu, cw, kHbeMaxStretch), pitch = __shfl_sync(0xffffffffu, cw, kHbePitch);
    const int usf4 = __shfl_sync(0xffffffffu, cw, kHbeUsf4);
    const int xo0 = __shfl_sync(0xffffffffu, cw, kHbeXover), xo1 = __shfl_sync(0xffffffffu, cw, kHbeXover + 1);
    const int xo2 = __shfl_sync(0xffffffffu, cw, kHbeXover + 2), xo3 = __shfl_sync(0xffffffffu, cw, kHbeXover + 3);
    int err = 0;
    if (usf4 || !(S == 4 || S == 8 || S == 12 || S == 16 || S == 20) || pitch < 0 || pitch > 127) err = -2;
    else if (ks < 0) err = -1;  // polyphase.c:187
    else if (ks + S > 32 || sb < 0 || eb > 64 || sb > eb) err = -2;
    else if (xo0 < 0 || xo1 > 64 || xo2 > 64 || xo3 > 64) err = -2;
    else if (ms >= 4 && xo2 <= 1) err = (int)0x80000000;  // hbe_trans.c:1572
    else if (ms >= 2 && (xo1 < xo0 || min(xo1, 63) - xo0 + 1 > 36)) err = -2;
    else if (ms >= 3 && xo2 < xo1) err = -2;
    else if (ms >= 4 && (xo3 < xo2 || min(63, ((xo3 - 1) >> 1) + 1) - max(0, (xo2 >> 1) - 1) + 1 > 18)) err = -2;
    if (err) {
      if (p.err && lane == 0) p.err[u] = err;
      continue;
    }
    float *st = p.state + u * kHbeStWords;
    const float *qre = p.qmf_re + u * p.in_stride, *qim = p.qmf_im + u * p.in_stride;
    float *pvr = p.pv_re + u * p.out_stride, *pvi = p.pv_im + u * p.out_stride;
    if (p.shift_rows) {  // ph_vocod_qmf rows 32..39 -> 0..7 (pvr / pvi point at row 8)
      float vr[16], vi[16];
#pragma unroll
      for (int q = 0; q < 16; q++) {
        vr[q] = pvr[64 * 24 + 32 * q + lane];
        vi[q] = pvi[64 * 24 + 32 * q + lane];
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 16; q++) {
        pvr[-512 + 32 * q + lane] = vr[q];
        pvi[-512 + 32 * q + lane] = vi[q];
      }
      __syncwarp();
    }
    const int A = 2 * S, e0 = 18 * S - 1;  // E[e0 + pos] = time sample `pos` of the reference's ptr_input_buf (pos >= 1)
    const bool fresh = S == 20;  // the reference re-initialises (clears both bank histories) on every call for this size

    // ---- phase 0: time-signal history ----
    for (int m = lane; m < 18 * S; m += 32) E[e0 - m] = fresh ? 0.0f : st[kHbeStAnal + m];  // x[m] of the analysis bank
    __syncwarp();
    for (int m = 1 + lane; m < S; m += 32) E[e0 + m] = st[kHbeStTail + m];
    // ---- phase A: real synthesis bank ----
    float *V = R1, *LOC = R1 + 41 * 41;
    const int VS = 2 * S + 1;
    {
      const float *ct = tabs + kHtCosTrans + ks * 32;
      for (int e4 = lane; e4 < 32 * S; e4 += 128) {  // four requests per array in flight (the loads head a dependent chain)
        float re[4], im[4];
        int kq[4], cq[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int e = min(e4 + 32 * q, 32 * S - 1);
          cq[q] = e / S;
          kq[q] = e - cq[q] * S;
          re[q] = __ldg(qre + cq[q] * 64 + ks + kq[q]);
          im[q] = __ldg(qim + cq[q] * 64 + ks + kq[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (e4 + 32 * q < 32 * S) LOC[kq[q] * 33 + cq[q]] = ct[(kq[q] << 1) + 0] * re[q] + ct[(kq[q] << 1) + 1] * im[q];
      }
      for (int e = lane; e < 9 * 2 * S; e += 32) {  // columns -1..-9 = synth_buf[0..18 S)
        const int m = e / (2 * S), pos = e - m * 2 * S;
        V[(8 - m) * VS + pos] = fresh ? 0.0f : st[kHbeStSynth + e];
      }
    }
    __syncwarp();
    {
      float *v = V + (lane + 9) * VS;
      if (S == 20) {  // polyphase.c:203-229; rows 0..9 of the table only produce values the later rows overwrite
        const float *pt = tabs + kHtSyn20 + 10 * 20;
        float loc[20];
#pragma unroll
        for (int k = 0; k < 20; k++) loc[k] = LOC[k * 33 + lane];
#pragma unroll 1
        for (int l = 10; l <= 30; l++) {
          float accu = 0.0f;
#pragma unroll
          for (int k = 0; k < 20; k++) accu += loc[k] * pt[k];
          pt += 20;
          if (l <= 20) { v[l] = accu; v[20 - l] = accu; }
          else if (l < 30) { v[l] = accu; v[60 - l] = -accu; }
          else v[30] = accu;
        }
      } else {
        float x[32], so[128];
        for (int k = 0; k < S; k++) { x[k] = LOC[k * 33 + lane]; x[k + S] = 0; }
        if (S == 12) real_synth_fft_p3(rom, x, so);
        else real_synth_fft_p2(rom + kHromFftTw, x, so, 2 * S);
        const float *pc = rom + kHromSynCos + syncos_off(S);
        const int k1 = S + (S >> 1);
        for (int k = 0; k < 2 * S; k++) {
          float tmp = so[2 * k] * __ldg(pc + 2 * k);
          tmp -= so[2 * k + 1] * __ldg(pc + 2 * k + 1);
          v[k < k1 ? (S >> 1) + k : k - k1] = tmp;
        }
      }
    }
    __syncwarp();
    {
      const float *win = tabs + kHtWin + win_off(S);
#pragma unroll 1
      for (int s = 0; s < S; s++) {
        float accu = 0.0f;
#pragma unroll
        for (int j = 0; j < 10; j++) {
          const int c = lane + 9 - (j >> 1) * 2 - (j & 1);
          accu = accu + V[c * VS + (j & 1) * S + s] * win[S * j + s];
        }
        E[e0 + (lane + 1) * S + s] = accu;
      }
    }
    __syncwarp();
    for (int e = lane; e < 18 * S; e += 32) {  // synth_buf[0..18 S) = columns 31..23
      const int m = e / (2 * S), pos = e - m * 2 * S;
      st[kHbeStSynth + e] = V[(40 - m) * VS + pos];
    }
    for (int m = lane; m < S; m += 32) st[kHbeStTail + m] = E[e0 + 32 * S + m];
    for (int m = lane; m < 18 * S; m += 32) st[kHbeStAnal + m] = E[e0 + 32 * S - m];
    __syncwarp();

    // ---- phase B: complex analysis bank, 16 columns ----
    float *U = R1;
    const int lo = 4 * ks, nq = 4 * S;
    {
      const float *win = tabs + kHtWin + win_off(A);
      for (int i = lane; i < 2 * A; i += 32) {  // lane = window phase: its five taps stay in registers for all 16 columns
        float wj[5];
#pragma unroll
        for (int j = 0; j < 5; j++) wj[j] = win[i + 2 * A * j];
        const float *ep = E + e0 + A - i;
#pragma unroll 4
        for (int col = 0; col < 16; col++) {
          float accu = 0.0f;
#pragma unroll
          for (int j = 0; j < 5; j++) accu = accu + ep[col * A - 2 * A * j] * wj[j];
          U[i * 17 + col] = accu;
        }
      }
      for (int e4 = lane; e4 < 12 * nq; e4 += 128) {  // rows 0..11 = the previous call's rows 16..27
        float v[4];
        int rq[4], cq[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int e = min(e4 + 32 * q, 12 * nq - 1);
          rq[q] = e / nq;
          cq[q] = e - rq[q] * nq;
          v[q] = st[kHbeStQin + rq[q] * 128 + lo + cq[q]];
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (e4 + 32 * q < 12 * nq) QM[rq[q] * kHbQW + cq[q]] = v[q];
      }
    }
    __syncwarp();
    if (S == 20) {  // polyphase.c:109-130
      const int col = lane & 15, half = lane >> 4;
      for (int i = half ? 20 : 1; i < (half ? 40 : 20); i++) {
        const float a = U[i * 17 + col], b = U[(80 - i) * 17 + col];
        U[i * 17 + col] = a + b;
        U[(80 - i) * 17 + col] = a - b;
      }
      __syncwarp();
      const float u0 = U[col], u40 = U[40 * 17 + col];
      float *row = QM + (col + 12) * kHbQW;
#pragma unroll 1
      for (int kk = 0; kk < 5; kk++) {  // four bins per pass share the sample loads
        const int k = half * 20 + 4 * kk;
        const float *t0 = tabs + kHtAna40 + k * 82;
        float ar[4], ai[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          ar[q] = u40;
          ai[q] = (q & 1) ? u0 : -u0;  // k is even
        }
#pragma unroll 3
        for (int l = 1; l < 40; l++) {
          const float ul = U[l * 17 + col], um = U[(80 - l) * 17 + col];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const float2 w = *reinterpret_cast<const float2 *>(t0 + 82 * q + 2 * l);
            ar[q] = ar[q] + ul * w.x;
            ai[q] = ai[q] + um * w.y;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          row[2 * (k + q)] = ar[q];
          row[2 * (k + q) + 1] = ai[q];
        }
      }
    } else if (lane < 16) {
      float u_in[256], u_out[256];
      const float *cs = rom + kHromAnaCs + anacs_off(S);
      for (int k = 0; k < 2 * A; k++) {
        const float uk = U[k * 17 + lane];
        u_in[2 * k] = __ldg(cs + 2 * k) * uk;
        u_in[2 * k + 1] = __ldg(cs + 2 * k + 1) * uk;
      }
      if (S == 12) cmplx_anal_fft_p3(rom, u_in, u_out);
      else cmplx_anal_fft_p2(rom + kHromFftTw, u_in, u_out, 2 * A);
      float *ab = QM + (lane + 12) * kHbQW;
      for (int k = 0; k < S; k++) {
        ab[4 * k + 1] = -u_out[4 * k];
        ab[4 * k] = u_out[4 * k + 1];
        ab[4 * k + 3] = u_out[4 * k + 2];
        ab[4 * k + 2] = -u_out[4 * k + 3];
      }
    }
    __syncwarp();
    for (int e = lane; e < 12 * nq; e += 32) {  // the next call's rows 0..11
      const int r = e / nq, c = e - r * nq;
      st[kHbeStQin + r * 128 + lo + c] = QM[(16 + r) * kHbQW + c];
    }

    // ---- phase C: products ----
    const QinView Q = {QM, lo, nq};
    const float pf = (float)((double)pitch * 0.08333333333333);  // hbe_trans.c:1557, 2:1 system
    const bool xp = !(pf < 1.0f);
    float *N2 = R1, *N4 = E, *OUTC = R1 + kNRows * kNW;
    static_assert(kNRows * kNW + 42 * kOW <= kHbR1 && kNRows * kN4W <= kHbE, "phase C buffers");
    const int n4lo = max(0, (xo2 >> 1) - 1), n4hi = min(63, ((xo3 - 1) >> 1) + 1);
    if (ms >= 2) {
      const int nb = min(xo1, 63) - xo0 + 1;
      for (int e = lane; e < 25 * nb; e += 32) {  // rows 1..25 are the ones the products read
        const int r = 1 + e / nb, b = e - (r - 1) * nb;
        const float xr = Q.at(r, 2 * (xo0 + b)), xi = Q.at(r, 2 * (xo0 + b) + 1);
        const float m = mag2(xr, xi);
        N2[r * kNW + 2 * b] = xr * m;
        N2[r * kNW + 2 * b + 1] = xi * m;
      }
    }
    if (ms >= 4) {
      const int nb = n4hi - n4lo + 1;
      for (int e = lane; e < kNRows * nb; e += 32) {
        const int r = e / nb, b = e - r * nb;
        const float xr = Q.at(r, 2 * (n4lo + b)), xi = Q.at(r, 2 * (n4lo + b) + 1);
        const float m = mag4(xr, xi);
        N4[r * kN4W + 2 * b] = xr * m;
        N4[r * kN4W + 2 * b + 1] = xi * m;
      }
    }
    __syncwarp();
    const int ci = lane & 15, bp = lane >> 4;
    for (int b0 = sb; b0 < eb; b0 += 8) {
      // the chunk is planar: column = 8 * (0 re | 1 im) + band - b0, so that a half-warp's two band parities fall on odd / even banks
      {
        float v[5];
#pragma unroll
        for (int q = 0; q < 5; q++) {  // carry rows 0..9
          const int e = lane + 32 * q, r = e >> 4, c = e & 15, band = b0 + (c & 7);
          v[q] = band < eb ? st[kHbeStQout + r * 128 + 2 * band + (c >> 3)] : 0.0f;
        }
#pragma unroll
        for (int q = 0; q < 5; q++) {
          const int e = lane + 32 * q;
          OUTC[(e >> 4) * kOW + (e & 15)] = v[q];
        }
      }
      for (int e = 160 + lane; e < 42 * 16; e += 32) OUTC[(e >> 4) * kOW + (e & 15)] = 0.0f;
      __syncwarp();
      // bands of this chunk that are neither stretch-3 nor stretch-4 (the common case: max_stretch = 2): the four bands a lane
      // owns walk the taps together, one lock step per tap instead of four
      const bool only2 = !((ms >= 3 && b0 + 8 > xo1 && b0 < xo2) || (ms >= 4 && b0 + 8 > xo2 && b0 < xo3));
      if (only2) {
        float xr4[4], xi4[4];
        XAdd xa4[4];
        bool act[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int b = b0 + 2 * q + bp;
          act[q] = b < eb && ms >= 2 && b >= xo0 && b < xo1;
          xa4[q] = XAdd{0, 0, 0, 0, false};
          xr4[q] = xi4[q] = 0;
          if (act[q]) {
            xr4[q] = N2[(6 + ci) * kNW + 2 * (b - xo0)];
            xi4[q] = N2[(6 + ci) * kNW + 2 * (b - xo0) + 1];
            if (xp) xa4[q] = xprod2(Q, N2, xo0, min(xo1, 63), b, ci, pf, rom + kHromXp2 + (pitch << 1));
          }
        }
#pragma unroll 1
        for (int k = 9; k >= 0; k--) {
          const float *nrow = N2 + (1 + ci + k) * kNW + 2 * (b0 + bp - xo0);
          float *orow = OUTC + (1 + 2 * ci + k) * kOW + bp;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            if (act[q]) {
              const float tr = nrow[4 * q], ti = nrow[4 * q + 1];
              const float cr = (tr * xr4[q] - ti * xi4[q]) * 0.3333333f;
              const float cim = (tr * xi4[q] + ti * xr4[q]) * 0.3333333f;
              float o0 = orow[2 * q] + cr, o1 = orow[8 + 2 * q] + cim;
              if (xa4[q].on && k == 4) { o0 += xa4[q].r0; o1 += xa4[q].i0; }
              if (xa4[q].on && k == 5) { o0 += xa4[q].r1; o1 += xa4[q].i1; }
              orow[2 * q] = o0;
              orow[8 + 2 * q] = o1;
            }
          }
          __syncwarp();
        }
      } else {
#pragma unroll 1
      for (int q = 0; q < 4; q++) {
        const int bb = 2 * q + bp, b = b0 + bb;
        int T = 0;
        if (b < eb) {
          if (ms >= 2 && b >= xo0 && b < xo1) T = 2;
          else if (ms >= 3 && b >= xo1 && b < xo2) T = 3;
          else if (ms >= 4 && b >= xo2 && b < xo3) T = 4;
        }
        float xzr = 0, xzi = 0, yr = 0, yi = 0, vx[16], vc[16];
        bool two = false;
        int ip = 0;
        XAdd xa = {0, 0, 0, 0, false};
        if (T == 2) {
          xzr = N2[(6 + ci) * kNW + 2 * (b - xo0)];
          xzi = N2[(6 + ci) * kNW + 2 * (b - xo0) + 1];
          if (xp) xa = xprod2(Q, N2, xo0, min(xo1, 63), b, ci, pf, rom + kHromXp2 + (pitch << 1));
        } else if (T == 3) {
          prod3_vectors(QM, lo, nq, rom, b, ci, vx, vc, two);
          if (!two) {
            const float tr = vx[8], ti = vx[9];
            xzr = tr * tr - ti * ti;
            xzi = tr * ti + ti * tr;
          } else {
            const float tr = vc[8], ti = vc[9], tr1 = vx[8], ti1 = vx[9];
            xzr = tr * tr - ti * ti;
            xzi = tr * ti + ti * tr;
            yr = tr1 * tr1 - ti1 * ti1;
            yi = tr1 * ti1 + ti1 * tr1;
          }
          if (xp) xa = xprod3(QM, lo, nq, rom, b, ci, pf, pitch);
        } else if (T == 4) {
          const int inp = b >> 1;
          ip = (b & 1) ? (inp + 1) : (inp - 1);
          float xr = N4[(6 + ci) * kN4W + 2 * (inp - n4lo)], xi = N4[(6 + ci) * kN4W + 2 * (inp - n4lo) + 1];
          const float tr = xr, ti = xi;
          float t = xr * xr - xi * xi;
          xi = xr * xi + xi * xr;
          xzr = tr * t - ti * xi;
          xzi = tr * xi + ti * t;
          if (xp) xa = xprod4(QM, lo, nq, rom, b, ci, pf, pitch);
        }
        const int klen = T == 2 ? 10 : T == 3 ? 8 : T == 4 ? 6 : 0;
        const int kx0 = 6 - T;  // the tap whose row is 2 col + 5
#pragma unroll 1
        for (int k = 9; k >= 0; k--) {
          if (k < klen) {
            float cr, cim;
            if (T == 2) {
              const float tr = N2[(1 + ci + k) * kNW + 2 * (b - xo0)], ti = N2[(1 + ci + k) * kNW + 2 * (b - xo0) + 1];
              cr = (tr * xzr - ti * xzi) * 0.3333333f;
              cim = (tr * xzi + ti * xzr) * 0.3333333f;
            } else if (T == 3) {
              float ar = vx[2 * k] * xzr - vx[2 * k + 1] * xzi, ai = vx[2 * k] * xzi + vx[2 * k + 1] * xzr;
              if (!two) {
                cr = ar * 0.4714045f;
                cim = ai * 0.4714045f;
              } else {
                ar += vc[2 * k] * yr - vc[2 * k + 1] * yi;
                ai += vc[2 * k] * yi + vc[2 * k + 1] * yr;
                cr = ar * 0.23570225f;
                cim = ai * 0.23570225f;
              }
            } else {
              const float a = N4[(ci + 2 * k) * kN4W + 2 * (ip - n4lo)], bi = N4[(ci + 2 * k) * kN4W + 2 * (ip - n4lo) + 1];
              cr = (a * xzr - bi * xzi) * 0.6666667f;
              cim = (a * xzi + bi * xzr) * 0.6666667f;
            }
            float *cell = OUTC + (T - 1 + 2 * ci + k) * kOW + bb;
            float o0 = cell[0] + cr, o1 = cell[8] + cim;
            if (xa.on && k == kx0) { o0 += xa.r0; o1 += xa.i0; }
            if (xa.on && k == kx0 + 1) { o0 += xa.r1; o1 += xa.i1; }
            cell[0] = o0;
            cell[8] = o1;
          }
          __syncwarp();
        }
      }
      }
      // ---- phase D: rotation (hbe_trans.c:280-294), carry rows ----
      for (int e = lane; e < 32 * 8; e += 32) {
        const int r = e >> 3, c = e & 7, band = b0 + c;
        if (band < eb) {
          const float a = OUTC[r * kOW + c], cc = OUTC[r * kOW + 8 + c];
          const float pc = __ldg(rom + kHromPvCos + band), ps = __ldg(rom + kHromPvSin + band);
          pvr[r * 64 + band] = a * pc - cc * ps;
          pvi[r * 64 + band] = a * ps + cc * pc;
        }
      }
      for (int e = lane; e < 10 * 16; e += 32) {
        const int r = e >> 4, c = e & 15, band = b0 + (c & 7);
        if (band < eb) st[kHbeStQout + r * 128 + 2 * band + (c >> 3)] = OUTC[(32 + r) * kOW + c];
      }
      __syncwarp();
    }
    if (p.err && lane == 0) p.err[u] = 0;
  }
}

cudaError_t launch_esbr_hbe(const EsbrHbeArgs &args, int num_sms, cudaStream_t stream) {
  static PerDeviceOnce configured;
  const size_t smem = (size_t)(kHtWords + kHbWarps * kHbWarpWords) * sizeof(float);
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(esbr_hbe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.done();
  }
  long long need = (args.n_units + kHbWarps - 1) / kHbWarps;
  long long grid = (long long)num_sms;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  esbr_hbe_kernel<<<(unsigned)grid, kHbWarps * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace xb
