/* esbr_hfgen.c — TEST INFRASTRUCTURE ONLY: plain-C restatement of the float eSBR HF generator
 *   ixheaacd_generate_hf                decoder/ixheaacd_sbrdec_lpfuncs.c:981-1359
 *   ixheaacd_esbr_calc_co_variance      :781-830      ixheaacd_esbr_chirp_fac_calc  :832-849
 *   ixheaacd_find_closest_entry         :263-285
 * for the 2:1 system without pre-processing, LD-MPS or error concealment.  Every float operation is written out one
 * rounding at a time in the reference's evaluation order (the reference build has no FMA and no reassociation), so the
 * results are bit-identical; pinned against the compiled function itself (tests/test_oracle_esbr.py). */
#include <math.h>
#include <string.h>
#include "xaac_oracle.h"

#define ROW(b, i) ((b) + 64 * ((i) + 2)) /* row i of the reference's offset pointer */

typedef struct {
  float p01r, p01i, p02r, p02i, p11, p12r, p12i, p22, det;
} cov_t;

static const float new_bw_table[4][4] = {/* lpfuncs.c:80 */
                                         {0.00f, 0.60f, 0.90f, 0.98f},
                                         {0.60f, 0.75f, 0.90f, 0.98f},
                                         {0.00f, 0.75f, 0.90f, 0.98f},
                                         {0.00f, 0.75f, 0.90f, 0.98f}};

/* lpfuncs.c:781 */
static void co_variance(cov_t *c, const float *re, const float *im, int bd, int len) {
  memset(c, 0, sizeof(*c));
  for (int j = 0; j < len; j++) {
    const float r0 = ROW(re, j)[bd], i0 = ROW(im, j)[bd];
    const float r1 = ROW(re, j - 1)[bd], i1 = ROW(im, j - 1)[bd];
    const float r2 = ROW(re, j - 2)[bd], i2 = ROW(im, j - 2)[bd];
    volatile float a, b; /* one rounding per operation */
    a = r0 * r1; b = i0 * i1; a = a + b; c->p01r = c->p01r + a;
    a = i0 * r1; b = r0 * i1; a = a - b; c->p01i = c->p01i + a;
    a = r0 * r2; b = i0 * i2; a = a + b; c->p02r = c->p02r + a;
    a = i0 * r2; b = r0 * i2; a = a - b; c->p02i = c->p02i + a;
    a = r1 * r1; b = i1 * i1; a = a + b; c->p11 = c->p11 + a;
    a = r1 * r2; b = i1 * i2; a = a + b; c->p12r = c->p12r + a;
    a = i1 * r2; b = r1 * i2; a = a - b; c->p12i = c->p12i + a;
    a = r2 * r2; b = i2 * i2; a = a + b; c->p22 = c->p22 + a;
  }
  {
    volatile float a, b, d;
    a = c->p11 * c->p22;
    b = c->p12r * c->p12r;
    d = c->p12i * c->p12i;
    b = b + d;
    b = b * 0.999999f;
    c->det = a - b;
  }
}

/* lpfuncs.c:1056-1098 (and 1264-1296): alpha[0..1] from the covariance */
static void solve_alpha(const cov_t *c, float ar[2], float ai[2]) {
  volatile float t, u, fac;
  if (c->det == 0.0f) {
    ar[1] = ai[1] = 0;
  } else {
    fac = 1.0f / c->det;
    t = c->p01r * c->p12r; u = c->p01i * c->p12i; t = t - u; u = c->p02r * c->p11; t = t - u;
    ar[1] = t * fac;
    t = c->p01i * c->p12r; u = c->p01r * c->p12i; t = t + u; u = c->p02i * c->p11; t = t - u;
    ai[1] = t * fac;
  }
  if (c->p11 == 0) {
    ar[0] = ai[0] = 0;
  } else {
    fac = 1.0f / c->p11;
    u = ar[1] * c->p12r; t = c->p01r + u; u = ai[1] * c->p12i; t = t + u;
    ar[0] = -t * fac;
    u = ai[1] * c->p12r; t = c->p01i + u; u = ar[1] * c->p12i; t = t - u;
    ai[0] = -t * fac;
  }
  {
    volatile float m0, m1;
    t = ar[0] * ar[0]; u = ai[0] * ai[0]; m0 = t + u;
    t = ar[1] * ar[1]; u = ai[1] * ai[1]; m1 = t + u;
    if (m0 >= 16.0f || m1 >= 16.0f) ar[0] = ai[0] = ar[1] = ai[1] = 0.0f;
  }
}

static int closest_entry_down(int goal, const int32_t *fm, int num_mf) { /* lpfuncs.c:263, direction 0 */
  if (goal <= fm[0]) return fm[0];
  if (goal >= fm[num_mf]) return fm[num_mf];
  int idx = num_mf;
  while (fm[idx] > goal) idx--;
  return fm[idx];
}

/* ixheaacd_gausssolve / ixheaacd_polyfit / ixheaacd_pre_processing (decoder/ixheaacd_sbrdec_lpfuncs.c:851-979): the spectral tilt of
 * the low band (third-order fit of its per-band level in dB) becomes a gain per source band of the patches.  Every float
 * operation is its own statement through a volatile so that no contraction can change a rounding; log10 / pow are libm's, in
 * double, as in the reference. */
static void pre_processing(const float *src_re, const float *src_im, float *gain, int n, int start, int end) {
  float low_env[64] = {0};
  volatile float mean = 0, t, u, v_;
  if (n != 0 && end != start) {
    for (int k = 0; k < n; k++) {
      volatile float temp = 0;
      for (int i = start; i < end; i++) {
        t = ROW(src_re, i)[k] * ROW(src_re, i)[k];
        u = ROW(src_im, i)[k] * ROW(src_im, i)[k];
        t = t + u;
        temp = temp + t;
      }
      temp = temp / (float)(end - start);
      t = temp + 1.0f;
      low_env[k] = (float)(10 * log10((double)t));
      mean = mean + low_env[k];
    }
    mean = mean / (float)n;
  }
  float a[4][4], b[4], vv[7], p[4];
  for (int i = 0; i < 4; i++) {
    b[i] = 0.0f;
    for (int j = 0; j < 4; j++) a[i][j] = 0.0f;
  }
  for (int k = 0; k < n; k++) {
    vv[0] = 1.0f;
    for (int i = 1; i <= 6; i++) {
      t = (float)k * vv[i - 1];
      vv[i] = t;
    }
    for (int i = 0; i <= 3; i++) {
      t = vv[3 - i] * low_env[k];
      u = b[i] + t;
      b[i] = u;
      for (int j = 0; j <= 3; j++) {
        u = a[i][j] + vv[6 - i - j];
        a[i][j] = u;
      }
    }
  }
  for (int i = 0; i < 4; i++) { /* :851 */
    int imax = i;
    for (int k = i + 1; k < 4; k++)
      if (fabs(a[k][i]) > fabs(a[imax][i])) imax = k;
    if (imax != i) {
      float w = b[imax];
      b[imax] = b[i];
      b[i] = w;
      for (int j = i; j < 4; j++) {
        w = a[imax][j];
        a[imax][j] = a[i][j];
        a[i][j] = w;
      }
    }
    v_ = a[i][i];
    t = b[i] / v_;
    b[i] = t;
    for (int j = i; j < 4; j++) {
      t = a[i][j] / v_;
      a[i][j] = t;
    }
    for (int k = i + 1; k < 4; k++) {
      v_ = a[k][i];
      t = v_ * b[i];
      u = b[k] - t;
      b[k] = u;
      for (int j = i + 1; j < 4; j++) {
        t = v_ * a[i][j];
        u = a[k][j] - t;
        a[k][j] = u;
      }
    }
  }
  for (int i = 3; i >= 0; i--) {
    p[i] = b[i];
    for (int j = i + 1; j < 4; j++) {
      t = a[i][j] * p[j];
      u = p[i] - t;
      p[i] = u;
    }
  }
  for (int k = 0; k < n; k++) {
    volatile float x = (float)k, sl = p[3];
    t = p[2] * x; sl = sl + t;
    x = x * x;
    t = p[1] * x; sl = sl + t;
    x = x * (float)k;
    t = p[0] * x; sl = sl + t;
    t = mean - sl;
    u = t / 20.0f;
    gain[k] = (float)pow(10, (double)u);
  }
}

int xo_esbr_generate_hf(const float *src_re, const float *src_im, const float *pv_re, const float *pv_im, float *dst_re,
                        float *dst_im, const int32_t *par, float *bw_prev, int32_t *patch_out) {
  const int32_t *fm = par + XO_EHF_FMASTER, *invf_tbl = par + XO_EHF_INVF_TBL;
  const int num_mf = par[XO_EHF_NUM_MF], num_if = par[XO_EHF_NUM_IF], sb_start = par[XO_EHF_SB_START];
  const int hbe = par[XO_EHF_HBE_FLAG], patching = par[XO_EHF_PATCHING_MODE];
  if (par[XO_EHF_USF4] || num_mf < 1 || num_mf > 56 || num_if < 0 || num_if > 5) return -2;
  const int lsb = fm[0], usb = fm[num_mf], xover = sb_start - fm[0];
  const int start = par[XO_EHF_BORDER_FIRST] * 2, end = 32 + (par[XO_EHF_BORDER_LAST] - 16) * 2, cov_len = 38;
  if (start < 0 || end > XO_EHF_ROWS - 2 || lsb < 0 || usb > 64 || lsb > usb) return -2;
  for (int i = 0; i < 5; i++)
    if ((par[XO_EHF_INVF + i] | par[XO_EHF_INVF_PREV + i]) & ~3) return -2;
  if (par[XO_EHF_FS] <= 0) return -2;
  float bw_array[6] = {0};
  int patch = 0;
  const int pre_proc = par[XO_EHF_PRE_PROC] != 0;
  float gain_vector[64];
  for (int k = 0; k < 64; k++) gain_vector[k] = 1.0f;
  if (pre_proc) pre_processing(src_re, src_im, gain_vector, lsb, start, end); /* lpfuncs.c:1052 */

  for (int i = 0; i < num_if; i++) { /* lpfuncs.c:832 */
    volatile float a, b;
    float bw = new_bw_table[par[XO_EHF_INVF_PREV + i]][par[XO_EHF_INVF + i]];
    if (bw < bw_prev[i]) {
      a = 0.75000f * bw; b = 0.25000f * bw_prev[i];
    } else {
      a = 0.90625f * bw; b = 0.09375f * bw_prev[i];
    }
    bw = a + b;
    if (bw < 0.015625) bw = 0;
    bw_array[i] = bw;
  }
  for (int i = start; i < end; i++)
    for (int k = usb; k < 64; k++) ROW(dst_re, i)[k] = ROW(dst_im, i)[k] = 0.0f;

  if (patching || !hbe) {
    int flag_break = 0;
    float ar[64][2], ai[64][2];
    memset(ar, 0, sizeof(ar));
    memset(ai, 0, sizeof(ai));
    int cov_count = lsb;
    if (par[XO_EHF_MPS_SBR]) cov_count = lsb < par[XO_EHF_COV_COUNT] ? lsb : par[XO_EHF_COV_COUNT];
    for (int k = 1; k < cov_count; k++) {
      cov_t c;
      co_variance(&c, src_re, src_im, k, cov_len);
      solve_alpha(&c, ar[k], ai[k]);
    }
    volatile float g = 2.048e6f / (float)par[XO_EHF_FS];
    g = g + 0.5f;
    int goal_sb = (int)g;
    if (goal_sb < fm[num_mf]) {
      int idx = 0;
      while (fm[idx] < goal_sb) idx++;
      goal_sb = fm[idx];
    } else {
      goal_sb = fm[num_mf];
    }
    int src_start = xover + 1, sb = lsb + xover;
    while (sb < usb) {
      if (patch >= 6) return -1;
      patch_out[1 + patch] = sb;
      int nb = goal_sb - sb, stride;
      if (nb >= lsb - src_start) {
        stride = (sb - src_start) & ~1;
        nb = lsb - (sb - stride);
        nb = closest_entry_down(sb + nb, fm, num_mf) - sb;
      }
      stride = (nb + sb - lsb + 1) & ~1;
      src_start = 1;
      if (goal_sb - (sb + nb) < 3) goal_sb = usb;
      if (nb < 3 && patch > 0 && sb + nb == usb) {
        for (int i = start; i < end; i++)
          for (int k2 = sb; k2 < sb + nb; k2++) ROW(dst_re, i)[k2] = ROW(dst_im, i)[k2] = 0.0f;
        break;
      }
      if (nb < 0 && flag_break == 1) break;
      if (nb < 0) {
        flag_break = 1;
        continue;
      }
      flag_break = 0;
      for (int k2 = sb; k2 < sb + nb; k2++) {
        const int k = k2 - stride;
        if (k < 0 || k >= 64) return -2; /* the reference would read outside the row */
        int bwi = 0;
        while (k2 >= invf_tbl[bwi]) {
          bwi++;
          if (bwi >= 5) return -1;
        }
        volatile float bw = bw_array[bwi], a0r, a0i, a1r, a1i;
        a0r = bw * ar[k][0];
        a0i = bw * ai[k][0];
        bw = bw * bw;
        a1r = bw * ar[k][1];
        a1i = bw * ai[k][1];
        for (int i = start; i < end; i++) {
          volatile float t, u, dr, di;
          const float gain = gain_vector[k];
          dr = ROW(src_re, i)[k] * gain;
          di = ROW(src_im, i)[k] * gain;
          if (bw > 0.0f) {
            const float r1 = ROW(src_re, i - 1)[k], i1 = ROW(src_im, i - 1)[k];
            const float r2 = ROW(src_re, i - 2)[k], i2 = ROW(src_im, i - 2)[k];
            t = a0r * r1; u = a0i * i1; t = t - u; u = a1r * r2; t = t + u; u = a1i * i2; t = t - u;
            t = t * gain;
            dr = dr + t;
            t = a0i * r1; u = a0r * i1; t = t + u; u = a1i * r2; t = t + u; u = a1r * i2; t = t + u;
            t = t * gain;
            di = di + t;
          }
          ROW(dst_re, i)[k2] = dr;
          ROW(dst_im, i)[k2] = di;
        }
      }
      sb += nb;
      patch++;
    }
  }

  if (pv_re && pv_im && hbe && !patching) {
    int bwi = 0;
    patch = 1;
    for (int k2 = sb_start; k2 < fm[num_mf]; k2++) {
      cov_t c;
      float ar[2], ai[2];
      co_variance(&c, pv_re, pv_im, k2, cov_len);
      solve_alpha(&c, ar, ai);
      while (k2 >= invf_tbl[bwi]) {
        bwi++;
        if (bwi >= 5) return -1;
      }
      volatile float bw = bw_array[bwi], a0r, a0i, a1r, a1i;
      a0r = bw * ar[0];
      a0i = bw * ai[0];
      bw = bw * bw;
      a1r = bw * ar[1];
      a1i = bw * ai[1];
      for (int i = start; i < end; i++) {
        float dr = ROW(pv_re, i)[k2], di = ROW(pv_im, i)[k2];
        if (bw > 0.0f) {
          const float r1 = ROW(pv_re, i - 1)[k2], i1 = ROW(pv_im, i - 1)[k2];
          const float r2 = ROW(pv_re, i - 2)[k2], i2 = ROW(pv_im, i - 2)[k2];
          volatile float t, u, v;
          t = a0r * r1; u = a0i * i1; t = t - u; u = a1r * r2; v = a1i * i2; u = u - v; t = t + u;
          dr = dr + t;
          t = a0i * r1; u = a0r * i1; t = t + u; u = a1i * r2; v = a1r * i2; u = u + v; t = t + u;
          di = di + t;
        }
        ROW(dst_re, i)[k2] = dr;
        ROW(dst_im, i)[k2] = di;
      }
    }
  }
  if (patch >= 7) return -1;
  patch_out[0] = patch;
  for (int i = 0; i < num_if; i++) bw_prev[i] = bw_array[i];
  return 0;
}

void xo_esbr_generate_hf_batch(const float *src_re, const float *src_im, const float *pv_re, const float *pv_im,
                               float *dst_re, float *dst_im, const int32_t *par, float *bw_prev, int32_t *patch_out,
                               int32_t *err, int n) {
  const size_t B = (size_t)XO_EHF_ROWS * 64;
  for (int u = 0; u < n; u++)
    err[u] = xo_esbr_generate_hf(src_re + u * B, src_im + u * B, pv_re ? pv_re + u * B : 0, pv_im ? pv_im + u * B : 0,
                                 dst_re + u * B, dst_im + u * B, par + (size_t)u * XO_EHF_PAR_WORDS, bw_prev + 6 * u,
                                 patch_out + 8 * u);
}
