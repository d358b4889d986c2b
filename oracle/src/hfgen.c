/*
 * oracle/src/hfgen.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain-C restatement of the fixed-point complex ("HQ") SBR HF generator of libxaac:
 * inverse-filter level emphasis, 2nd-order complex covariance per low band, LPC coefficient solve,
 * bandwidth expansion and patch construction.  Cites reference lines (paths relative to /root/reference).
 * Pinned against the compiled reference (ref_hf_generator_hq) and against records tapped from real decodes.
 */
#include <string.h>
#include "fixmath.h"
#include "xaac_oracle.h"

/* common/ixheaac_basic_ops32.h:134 — NOT commutative: the second operand contributes only its high half */
static inline i32 hm(i32 a, i32 b) { return (i32)(((i64)a * (i64)(b >> 16)) >> 16); }
static inline i32 abs_sat(i32 a) { return a == OX_MIN32 ? OX_MAX32 : (a < 0 ? -a : a); } /* ops32.h:293 */

/* decoder/ixheaacd_basic_funcs.c:130-152 */
static i32 fix_div(i32 op1, i32 op2) {
  i32 q = 0;
  u32 num = (u32)(((op1 >> 1) < 0) ? -(op1 >> 1) : (op1 >> 1));
  u32 den = (u32)(((op2 >> 1) < 0) ? -(op2 >> 1) : (op2 >> 1));
  if (num != 0)
    for (int k = 15; k > 0; k--) {
      q = (i32)((u32)q << 1);
      num <<= 1;
      if (num >= den) {
        num -= den;
        q++;
      }
    }
  return ((op1 ^ op2) < 0) ? -q : q;
}

/* ISO/IEC 14496-3 newBw table in Q31 (decoder/ixheaacd_sbrdec_lpfuncs.c:85-89): 0, 0.6, 0.75, 0.9, 0.98 */
static const i32 new_bw[4][4] = {{0x00000000, 0x4ccccccd, 0x73333333, 0x7d70a3d7},
                                 {0x4ccccccd, 0x60000000, 0x73333333, 0x7d70a3d7},
                                 {0x00000000, 0x60000000, 0x73333333, 0x7d70a3d7},
                                 {0x00000000, 0x60000000, 0x73333333, 0x7d70a3d7}};

typedef struct { i32 p11, p22, p01, p02, p12, p01i, p02i, p12i; } cov_t;

/* decoder/ixheaacd_lpp_tran.c:374-627, in closed form. X(n) = row n of band k, >>3; n = -2,-1 are the LPC states.
 *   phi_01 = sum_{m=0}^{L-1} A(m)   phi_12 = sum_{m=-1}^{L-2} A(m)   A(m) = hm(xr[m],xr[m-1]) + hm(xi[m],xi[m-1])
 *   (imag)   B(m) = hm(xi[m],xr[m-1]) - hm(xr[m],xi[m-1])
 *   phi_11 = sum_{m=-1}^{L-2} C(m)  phi_22 = sum_{m=-2}^{L-3} C(m)  C(m) = hm(xr[m],xr[m]) + hm(xi[m],xi[m])
 *   phi_02 = sum_{m=0}^{L-1} D(m)   D(m) = hm(xr[m],xr[m-2]) + hm(xi[m],xi[m-2]),  E(m) likewise for the imaginary part
 * all sums wrap (plain C adds under -fwrapv). */
static void covariance(const i32 *row_m2 /* row of n = -2 */, int k, int L, cov_t *c) {
#define XR(n) (row_m2[((n) + 2) * 128 + k] >> 3)
#define XI(n) (row_m2[((n) + 2) * 128 + 64 + k] >> 3)
  i32 p01 = 0, p12 = 0, p01i = 0, p12i = 0, p11 = 0, p22 = 0, p02 = 0, p02i = 0;
  for (int m = -1; m <= L - 1; m++) {
    i32 A = ox_add(hm(XR(m), XR(m - 1)), hm(XI(m), XI(m - 1)));
    i32 B = ox_sub(hm(XI(m), XR(m - 1)), hm(XR(m), XI(m - 1)));
    if (m >= 0) { p01 = ox_add(p01, A); p01i = ox_add(p01i, B); }
    if (m <= L - 2) { p12 = ox_add(p12, A); p12i = ox_add(p12i, B); }
  }
  for (int m = -2; m <= L - 2; m++) {
    i32 C = ox_add(hm(XR(m), XR(m)), hm(XI(m), XI(m)));
    if (m >= -1) p11 = ox_add(p11, C);
    if (m <= L - 3) p22 = ox_add(p22, C);
  }
  for (int m = 0; m <= L - 1; m++) {
    p02 = ox_add(p02, ox_add(hm(XR(m), XR(m - 2)), hm(XI(m), XI(m - 2))));
    p02i = ox_add(p02i, ox_sub(hm(XI(m), XR(m - 2)), hm(XR(m), XI(m - 2))));
  }
#undef XR
#undef XI
  c->p11 = p11; c->p22 = p22; c->p01 = p01; c->p02 = p02; c->p12 = p12; c->p01i = p01i; c->p02i = p02i; c->p12i = p12i;
}

/* decoder/ixheaacd_lpp_tran.c:956-1258.  lpc[2][128]: LPC states (rows n = -2, -1); matrix[38][128] in place. */
int xo_hf_generator_hq(const i32 *lpc, i32 *matrix, const i16 *prm, i32 *bw_prev) {
  const int num_patches = prm[XO_HF_NUM_PATCHES], start_patch = prm[XO_HF_START_PATCH];
  const int stop_patch = prm[XO_HF_STOP_PATCH], num_columns = prm[XO_HF_NUM_COLUMNS];
  const i16 *bw_borders = prm + XO_HF_BW_BORDERS;
  const i16 *patch = prm + XO_HF_PATCH;
  const int factor = prm[XO_HF_FACTOR], num_if_bands = prm[XO_HF_NUM_IF_BANDS];
  const int max_qmf_subband = prm[XO_HF_MAX_QMF_SUBBAND];
  const int auto_corr_len = (num_columns + 6 == 36) ? 36 : 38; /* :1034-1045 */
  int start_idx = prm[XO_HF_START_IDX] * factor;
  int stop_idx = num_columns + prm[XO_HF_STOP_IDX] * factor;
  i32 bw_array[6] = {0};
  int bw_index[6] = {0};
  /* scratch = [lpc row -2, lpc row -1, matrix rows 0..37]; only bands start_patch..stop_patch of the LPC rows are
   * defined in the reference's scratch (:1022-1030) — nothing else of them is ever read. */
  static __thread i32 x[40 * 128];
  memcpy(x, lpc, 2 * 128 * sizeof(i32));
  memcpy(x + 256, matrix, 38 * 128 * sizeof(i32));

  /* decoder/ixheaacd_sbrdec_lpfuncs.c:735-767 */
  for (int i = 0; i < num_if_bands; i++) {
    i32 nb = new_bw[prm[XO_HF_INVF_PREV + i]][prm[XO_HF_INVF + i]];
    i16 w1 = nb < bw_prev[i] ? 0x6000 : 0x7400, w2 = nb < bw_prev[i] ? 0x2000 : 0x0c00;
    i32 acc = ox_add(ox_mul32x16_shl(nb, w1), ox_mul32x16_shl(bw_prev[i], w2));
    if (acc < 0x02000000) acc = 0;
    if (acc >= 0x7f800000) acc = 0x7f800000;
    bw_array[i] = acc;
  }
  /* :995-1007 — clear everything above the last patch for the generated slots */
  int actual_stop = (i16)(patch[6 * (num_patches - 1) + 3] + patch[6 * (num_patches - 1) + 5]);
  for (int i = start_idx; i < stop_idx; i++)
    for (int b = actual_stop; b < 64; b++) x[256 + 128 * i + b] = x[256 + 128 * i + 64 + b] = 0;
  int common_scale = prm[XO_HF_OV_LB_SCALE] < prm[XO_HF_LB_SCALE] ? prm[XO_HF_OV_LB_SCALE] : prm[XO_HF_LB_SCALE];

  for (int lb = start_patch; lb < stop_patch; lb++) {
    cov_t s, c;
    covariance(x, lb, auto_corr_len, &s);
    int reset = 0;
    i32 mx = ox_abs_nrm(s.p01) | ox_abs_nrm(s.p02) | ox_abs_nrm(s.p12) | s.p11 | s.p22 | ox_abs_nrm(s.p01i) |
             ox_abs_nrm(s.p02i) | ox_abs_nrm(s.p12i);
    int q = ox_pnorm32(mx);
    c.p11 = ox_lsl(s.p11, q); c.p22 = ox_lsl(s.p22, q); c.p01 = ox_lsl(s.p01, q); c.p02 = ox_lsl(s.p02, q);
    c.p12 = ox_lsl(s.p12, q); c.p01i = ox_lsl(s.p01i, q); c.p02i = ox_lsl(s.p02i, q); c.p12i = ox_lsl(s.p12i, q);
    i32 m2 = ox_add_sat(ox_mul32(c.p12, c.p12), ox_mul32(c.p12i, c.p12i));
    i32 d = ox_shl1(ox_sub_sat(ox_mul32(c.p11, c.p22), m2));
    i16 ar[2] = {0, 0}, ai[2] = {0, 0};
    if (d != 0) { /* :1084-1126 */
      int nd = ox_norm32(d);
      i16 inv = (i16)fix_div(0x40000000, ox_lsl(d, nd));
      i32 mod_d = abs_sat(d);
      i32 tr = ox_sub_sat(ox_sub_sat(ox_mul32(c.p01, c.p12), ox_mul32(c.p01i, c.p12i)), ox_mul32(c.p02, c.p11)) >> 1;
      i32 ti = ox_sub_sat(ox_add_sat(ox_mul32(c.p01i, c.p12), ox_mul32(c.p01, c.p12i)), ox_mul32(c.p02i, c.p11)) >> 1;
      if (abs_sat(tr) >= mod_d) reset = 1;
      else ar[1] = (i16)(ox_lsl(ox_mul32x16(tr, inv), nd + 1) >> 15);
      if (abs_sat(ti) >= mod_d) reset = 1;
      else ai[1] = (i16)(ox_lsl(ox_mul32x16(ti, inv), nd + 1) >> 15);
    }
    if (c.p11 != 0) { /* :1131-1178 */
      int n11 = ox_norm32(c.p11);
      i16 inv = (i16)fix_div(0x40000000, ox_lsl(c.p11, n11));
      i32 tr = ox_add_sat(ox_add(c.p01 >> 3, ox_mul32x16(c.p12, ar[1])), ox_mul32x16(c.p12i, ai[1]));
      i32 ti = ox_sub_sat(ox_add(c.p01i >> 3, ox_mul32x16(c.p12, ai[1])), ox_mul32x16(c.p12i, ar[1]));
      tr = ox_shl1(tr);
      ti = ox_shl1(ti);
      if (abs_sat(tr) >= c.p11) reset = 1;
      else ar[0] = (i16)(ox_lsl(ox_mul32x16(ox_sub_sat(0, tr), inv), n11 + 1) >> 15);
      if (abs_sat(ti) >= c.p11) reset = 1;
      else ai[0] = (i16)(ox_lsl(ox_mul32x16(ox_sub_sat(0, ti), inv), n11 + 1) >> 15);
    }
    if (ox_add_sat((i32)ar[0] * ar[0], (i32)ai[0] * ai[0]) >= 0x40000000) reset = 1;
    if (ox_add_sat((i32)ar[1] * ar[1], (i32)ai[1] * ai[1]) >= 0x40000000) reset = 1;
    if (reset) ar[0] = ar[1] = ai[0] = ai[1] = 0;

    for (int p = 0; p < num_patches; p++) { /* :1200-1250 */
      const i16 *pp = patch + 6 * p;
      int hb = lb + pp[4];
      if (lb < pp[0] || lb >= pp[1]) continue;
      if (hb < max_qmf_subband) continue;
      while (bw_index[p] < 5 && bw_index[p] < 10 && hb >= bw_borders[bw_index[p]]) bw_index[p]++;
      i16 bw = (i16)(bw_array[bw_index[p]] >> 16);
      i16 a0r = ox_mult16_shl_sat(bw, ar[0]), a0i = ox_mult16_shl_sat(bw, ai[0]);
      bw = ox_mult16_shl_sat(bw, bw);
      i16 a1r = ox_mult16_shl_sat(bw, ar[1]), a1i = ox_mult16_shl_sat(bw, ai[1]);
      /* rows are relative to the scratch base (row 0 = LPC state n = -2): destination row t + 2 */
      for (int t = start_idx; t < stop_idx; t++) {
        const i32 *r2 = x + 128 * t, *r1 = x + 128 * (t + 1), *r0 = x + 128 * (t + 2);
        i32 *dst = x + 128 * (t + 2);
        if (bw > 0) { /* decoder/ixheaacd_lpp_tran.c:102-167 */
          i32 p2r = r2[lb], p2i = r2[64 + lb], p1r = r1[lb], p1i = r1[64 + lb];
          i32 acc = ox_sub(ox_add(ox_sub(ox_mul32x16(p1r, a0r), ox_mul32x16(p1i, a0i)), ox_mul32x16(p2r, a1r)),
                           ox_mul32x16(p2i, a1i));
          dst[hb] = ox_add(r0[lb] >> 2, ox_shl1(acc));
          acc = ox_add(ox_add(ox_add_sat(ox_add_sat(ox_mul32x16(p1r, a0i), ox_mul32x16(p1i, a0r)),
                                         ox_mul32x16(p2r, a1i)),
                              ox_mul32x16(p2i, a1r)), 0);
          dst[64 + hb] = ox_add(r0[64 + lb] >> 2, ox_shl1(acc));
        } else {
          dst[hb] = r0[lb] >> 2;
          dst[64 + hb] = r0[64 + lb] >> 2;
        }
      }
    }
  }
  memcpy(matrix, x + 256, 38 * 128 * sizeof(i32));
  for (int i = 0; i < num_if_bands; i++) bw_prev[i] = bw_array[i];
  return (i16)(common_scale - 2);
}

void xo_hf_generator_hq_batch(const i32 *lpc, i32 *matrix, const i16 *prm, i32 *bw_prev, i32 *hb_scale, int n) {
  for (int u = 0; u < n; u++)
    hb_scale[u] = xo_hf_generator_hq(lpc + (size_t)u * 256, matrix + (size_t)u * 38 * 128, prm + (size_t)u * XO_HF_PRM_WORDS,
                                     bw_prev + (size_t)u * 6);
}
