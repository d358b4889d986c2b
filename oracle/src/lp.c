/*
 * oracle/src/lp.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain-C restatement of the low-power (real-valued) fixed-point SBR path libxaac runs for stereo HE-AACv1
 * (low_pow_flag = 1, decoder/ixheaacd_sbrdecoder.c:408-419): the low-power HF generator with its alias-degree
 * estimation, and ixheaacd_sbr_dec's fixed branch for real matrices (64 words per slot).  Filterbanks are in qmf.c
 * (dct3_32 / dct2_64), envelope adjuster variants in envcalc_lp.inc.  Pinned against whole-stage records tapped from
 * a real HE-AACv1 stereo decode of the compiled reference (tests/golden/sbrdec_lp_tapped.npz).
 */
#include <string.h>
#include "fixmath.h"
#include "xaac_oracle.h"

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }
static inline i32 hm(i32 a, i32 b) { return (i32)(((i64)a * (i64)(b >> 16)) >> 16); } /* ops32.h:134 */
static inline i32 abs32(i32 a) { return a < 0 ? (i32)(0u - (u32)a) : a; }               /* ops32.h:271, wraps */
static inline i32 abs_sat(i32 a) { return a == OX_MIN32 ? OX_MAX32 : (a < 0 ? -a : a); }

/* decoder/ixheaacd_basic_funcs.c:130-152 */
static i32 fix_div(i32 op1, i32 op2) {
  i32 q = 0;
  u32 num = (u32)abs32(op1 >> 1), den = (u32)abs32(op2 >> 1);
  if (num != 0)
    for (int k = 15; k > 0; k--) {
      q = (i32)((u32)q << 1);
      num <<= 1;
      if (num >= den) { num -= den; q++; }
    }
  return ((op1 ^ op2) < 0) ? -q : q;
}

static const i32 new_bw[4][4] = {{0x00000000, 0x4ccccccd, 0x73333333, 0x7d70a3d7},
                                 {0x4ccccccd, 0x60000000, 0x73333333, 0x7d70a3d7},
                                 {0x00000000, 0x60000000, 0x73333333, 0x7d70a3d7},
                                 {0x00000000, 0x60000000, 0x73333333, 0x7d70a3d7}};

typedef struct { i32 p11, p22, p01, p02, p12, d; } cov_t;

/* decoder/ixheaacd_lpp_tran.c:271-372 for len = 38.  x: column of band k in the scratch (row stride 64, rows 0..39). */
static void covariance_lp(const i32 *x, cov_t *c) {
#define X(n) ox_shr32(x[64 * (n)], 3)
  i32 p01 = 0, p02 = 0, p11 = 0;
  for (int n = 2; n < 40; n++) {
    p01 = ox_add(p01, hm(X(n), X(n - 1)));
    p02 = ox_add(p02, hm(X(n), X(n - 2)));
    p11 = ox_add(p11, hm(X(n - 1), X(n - 1)));
  }
  i32 p12 = ox_add(ox_sub(p01, hm(X(39), X(38))), hm(X(1), X(0)));
  i32 p22 = ox_add(ox_sub(p11, hm(X(38), X(38))), hm(X(0), X(0)));
#undef X
  i32 mx = ox_abs_nrm(p01) | ox_abs_nrm(p02) | ox_abs_nrm(p12) | p11 | p22;
  int q = ox_pnorm32(mx);
  c->p11 = ox_lsl(p11, q); c->p22 = ox_lsl(p22, q); c->p01 = ox_lsl(p01, q); c->p02 = ox_lsl(p02, q);
  c->p12 = ox_lsl(p12, q);
  c->d = ox_sub_sat(ox_mul32(c->p22, c->p11), ox_mul32(c->p12, c->p12));
}

/* decoder/ixheaacd_lpp_tran.c:843-954 (+ filter1_lp :665-833, filt_step3_lp :629-663, invfilt_level_emphasis).
 * x: scratch of 40 rows x 64 words: rows 0,1 = LPC states (filled here), rows 2.. = matrix rows 0..37.
 * prm: XO_HF_* record; START_IDX / STOP_IDX are the already multiplied slot offsets the reference passes
 * (border[0] * time_step, time_step * (border[num_env] - num_time_slots)). */
void xo_low_pow_hf_generator(const i32 *lpc /* [2][128], 32 real words used */, i32 *x, const i16 *prm, i32 *bw_prev,
                             i16 *degree_alias /* [64] */, int norm_max) {
  const int num_patches = prm[XO_HF_NUM_PATCHES], num_columns = prm[XO_HF_NUM_COLUMNS];
  const i16 *patch = prm + XO_HF_PATCH, *bw_borders = prm + XO_HF_BW_BORDERS;
  const int num_if_bands = prm[XO_HF_NUM_IF_BANDS], max_qmf_subband = prm[XO_HF_MAX_QMF_SUBBAND];
  const int start_idx = prm[XO_HF_START_IDX];
  const int stop_idx = num_columns + prm[XO_HF_STOP_IDX];
  i32 *m = x + 128; /* matrix row 0 */
  i32 bw_array[6] = {0};
  cov_t cov[32];
  memset(cov, 0, sizeof(cov));
  for (int i = 0; i < num_if_bands; i++) { /* sbrdec_lpfuncs.c:735-767 */
    i32 nb = new_bw[prm[XO_HF_INVF_PREV + i]][prm[XO_HF_INVF + i]];
    i16 w1 = nb < bw_prev[i] ? 0x6000 : 0x7400, w2 = nb < bw_prev[i] ? 0x2000 : 0x0c00;
    i32 acc = ox_add(ox_mul32x16_shl(nb, w1), ox_mul32x16_shl(bw_prev[i], w2));
    if (acc < 0x02000000) acc = 0;
    if (acc >= 0x7f800000) acc = 0x7f800000;
    bw_array[i] = acc;
  }
  int actual_stop = ox_add16(patch[6 * (num_patches - 1) + 3], patch[6 * (num_patches - 1) + 5]);
  { /* :867-885 */
    int len = 6;
    if (len > stop_idx) len = stop_idx;
    for (int l = start_idx; l < len; l++)
      for (int b = actual_stop; b < 64; b++) m[64 * l + b] = 0;
    if (actual_stop < 32)
      for (int l = len; l < stop_idx; l++)
        for (int b = actual_stop; b < 32; b++) m[64 * l + b] = 0;
  }
  int start_patch = imax(1, ox_sub16(prm[XO_HF_START_PATCH], 2));
  int stop_patch = patch[3]; /* patch_param[0].dst_start_band */
  for (int i = 0; i < 2; i++) memcpy(x + 64 * i, lpc + 128 * i, sizeof(i32) * stop_patch);
  if (norm_max != 30)
    for (int k = start_patch; k < stop_patch; k++) covariance_lp(x + k, &cov[k]);

  { /* filter1_lp */
    i16 k1, k1_below = 0, k1_below2 = 0;
    int bw_index[6] = {0};
    for (int lb = start_patch; lb < stop_patch; lb++) {
      const cov_t *c = &cov[lb];
      i16 alpha[2] = {0, 0};
      if (c->d != 0) {
        int nd = ox_norm32(c->d);
        i16 inv = (i16)fix_div(0x40000000, ox_lsl(c->d, nd));
        i32 mod_d = abs32(c->d);
        i32 t = ox_sub_sat(ox_mul32(c->p01, c->p12), ox_mul32(c->p02, c->p11)) >> 2;
        if (abs32(t) < mod_d) {
          i32 v = (t == OX_MIN32 && inv == (i16)0x8000) ? OX_MAX32 : ox_mul32x16_shl(t, inv);
          alpha[1] = (i16)(ox_lsl(v, nd) >> 15);
        }
        t = ox_sub_sat(ox_mul32(c->p02, c->p12), ox_mul32(c->p01, c->p22)) >> 2;
        if (abs32(t) < mod_d) {
          i32 v = (t == OX_MIN32 && inv == (i16)0x8000) ? OX_MAX32 : ox_mul32x16_shl(t, inv);
          alpha[0] = (i16)(ox_lsl(v, nd) >> 15);
        }
      }
      if (c->p11 == 0) k1 = 0;
      else if (abs_sat(c->p01) >= c->p11) k1 = c->p01 < 0 ? 0x7fff : (i16)-0x8000;
      else k1 = (i16)(-(i16)fix_div(c->p01, c->p11));
      if (lb > 1) {
        i16 deg = ox_sub16_sat(0x7fff, ox_mult16_shl_sat(k1_below, k1_below));
        degree_alias[lb] = 0;
        if (((lb & 1) == 0) && k1 < 0) {
          if (k1_below < 0) {
            degree_alias[lb] = 0x7fff;
            if (k1_below2 > 0) degree_alias[lb - 1] = deg;
          } else if (k1_below2 > 0) degree_alias[lb] = deg;
        }
        if (((lb & 1) != 0) && k1 > 0) {
          if (k1_below > 0) {
            degree_alias[lb] = 0x7fff;
            if (k1_below2 < 0) degree_alias[lb - 1] = deg;
          } else if (k1_below2 < 0) degree_alias[lb] = deg;
        }
      }
      k1_below2 = k1_below;
      k1_below = k1;
      for (int p = 0; p < num_patches; p++) {
        const i16 *pp = patch + 6 * p;
        int hb = lb + pp[4];
        if (lb < pp[0] || lb >= pp[1] || hb < max_qmf_subband) continue;
        while (hb >= bw_borders[bw_index[p]]) bw_index[p]++;
        i32 bw_vec = bw_array[bw_index[p]];
        i16 bw = (i16)(bw_vec >> 16);
        i32 a0r = ox_shl32(ox_mult16x16(bw, alpha[0]), 1);
        bw = ox_mult16_shl_sat(bw, bw);
        i32 a1r = ox_shl32(ox_mult16x16(bw, alpha[1]), 1);
        const i32 *lo = x + lb + 64 * start_idx; /* scratch rows start_idx.. (row 0 = LPC state n = -2) */
        i32 *hi = x + hb + 64 * (start_idx + 2);
        int len = stop_idx - start_idx - 1;
        if (bw > 0) { /* :629-663 */
          i32 prev2 = lo[0], prev1 = lo[64];
          lo += 128;
          for (int i = len; i >= 0; i -= 2) {
            i32 curr = lo[0];
            i32 t = hm(prev2, a1r);
            lo += 64;
            hi[0] = ox_add_sat(curr >> 2, ox_shl1(ox_add(t, hm(prev1, a0r))));
            hi += 64;
            prev2 = lo[0];
            t = hm(prev1, a1r);
            lo += 64;
            hi[0] = ox_add_sat(prev2 >> 2, ox_shl1(ox_add(t, hm(curr, a0r))));
            hi += 64;
            prev1 = prev2;
            prev2 = curr;
          }
        } else {
          lo += 128;
          for (int i = len; i >= 0; i--, lo += 64, hi += 64) hi[0] = lo[0] >> 2;
        }
      }
    }
  }
  for (int lb = prm[XO_HF_START_PATCH]; lb < prm[XO_HF_STOP_PATCH]; lb++) /* :927-951 */
    for (int p = 0; p < num_patches; p++) {
      const i16 *pp = patch + 6 * p;
      int hb = lb + pp[4];
      if (lb < pp[0] || lb >= pp[1] || hb >= 64) continue;
      if (hb != pp[3]) degree_alias[hb] = degree_alias[lb];
    }
  for (int i = 0; i < num_if_bands; i++) bw_prev[i] = bw_array[i];
}

/* decoder/ixheaacd_sbrdec_lpfuncs.c:453-527 (real) */
static void rescale_x_overlap_lp(i32 *m, i16 *sf, i16 *misc, const i16 *env, int syn_usb) {
  int old_lsb = misc[XO_SBR_MISC_MAX_QMF_PREV];
  int start_slot = env[XO_ENV_TIME_STEP] * (misc[XO_SBR_MISC_END_POS_PREV] - env[XO_ENV_NUM_TIME_SLOTS]);
  int new_lsb = env[XO_ENV_MAX_QMF_SUBBAND];
  misc[XO_SBR_MISC_CODEC_USB] = (i16)new_lsb;
  misc[XO_SBR_MISC_SYN_LSB] = (i16)new_lsb;
  int b0 = imin(old_lsb, new_lsb), b1 = imax(old_lsb, new_lsb);
  if (new_lsb == old_lsb || old_lsb <= 0) return;
  for (int l = start_slot; l < 6; l++)
    for (int k = old_lsb; k < new_lsb; k++) m[64 * l + k] = 0;
  int source_scale, target_scale, t_lsb, t_usb;
  if (new_lsb > old_lsb) {
    source_scale = sf[XO_SF_OV_HB]; target_scale = sf[XO_SF_OV_LB]; t_lsb = 0; t_usb = old_lsb;
  } else {
    source_scale = sf[XO_SF_OV_LB]; target_scale = sf[XO_SF_OV_HB]; t_lsb = old_lsb; t_usb = syn_usb;
  }
  int reserve = xo_expsubbandsamples_lp(m, b0, b1, 0, start_slot);
  xo_adjust_scale_lp(m, b0, b1, 0, start_slot, reserve);
  source_scale += reserve;
  int delta = target_scale - source_scale;
  if (delta > 0) {
    delta = -delta;
    b0 = t_lsb;
    b1 = t_usb;
    if (new_lsb > old_lsb) sf[XO_SF_OV_LB] = (i16)source_scale;
    else sf[XO_SF_OV_HB] = (i16)source_scale;
  }
  xo_adjust_scale_lp(m, b0, b1, 0, start_slot, delta);
}

/* ixheaacd_sbr_dec, fixed branch with low_pow_flag = 1 (decoder/ixheaacd_sbr_dec.c:662-1310).  Same records as
 * xo_sbr_dec_hq; the overlap slots of the state blob hold 6 real slots of 64 words, the LPC rows 32 real words each.
 * scratch: 40 rows x 64 WORD32. */
int xo_sbr_dec_lp(const uint8_t *qrom, const uint8_t *env_rom, const uint8_t *misc_rom, const i16 *side, i16 *st,
                  const i16 *time_in, int ch_in, i16 *time_out, int ch_out, i32 *scratch) {
  const i16 *env = side + XO_SIDE_ENV, *hfs = side + XO_SIDE_HF;
  const int apply = side[XO_SIDE_APPLY];
  i16 *sf = st + XO_SBR_ST_SF, *misc = st + XO_SBR_ST_MISC;
  i32 *lpc = (i32 *)(st + XO_SBR_ST_LPC), *ov = (i32 *)(st + XO_SBR_ST_OV), *bw_prev = (i32 *)(st + XO_SBR_ST_BW_PREV);
  i32 *m = scratch + 128;
  const i16 *border = env + XO_ENV_BORDER_VEC;
  const int num_env = env[XO_ENV_NUM_ENV];
  memcpy(m, ov, 6 * 64 * sizeof(i32));
  sf[XO_SF_LB] = 0;
  if (apply) rescale_x_overlap_lp(m, sf, misc, env, misc[XO_SBR_MISC_SYN_USB]);
  {
    i32 pos = st[XO_SBR_ST_ANAL_POS], fpos = st[XO_SBR_ST_ANAL_POS + 1];
    sf[XO_SF_ST_LB] = 0;
    sf[XO_SF_LB] = (i16)xo_anal_qmffilt_lp(qrom, time_in, ch_in, st + XO_SBR_ST_ANAL_STATES, &pos, &fpos, m + 6 * 64);
    st[XO_SBR_ST_ANAL_POS] = (i16)pos;
    st[XO_SBR_ST_ANAL_POS + 1] = (i16)fpos;
  }
  int save_lb_scale, max_samp_val;
  {
    int usb = misc[XO_SBR_MISC_CODEC_USB];
    int reserve = xo_expsubbandsamples_lp(m, 0, usb, 6, 38);
    int reserve_ov1 = xo_expsubbandsamples_lp(m, 0, usb, 0, 6);
    max_samp_val = imin(reserve, reserve_ov1);
    i32 lrows[2 * 64];
    for (int i = 0; i < 2; i++) memcpy(lrows + 64 * i, lpc + 128 * i, 32 * sizeof(i32));
    int reserve_ov2 = xo_expsubbandsamples_lp(lrows, 0, usb, 0, 2);
    reserve_ov1 = imin(reserve_ov1, reserve_ov2);
    int shift1 = sf[XO_SF_LB] + reserve, shift2 = sf[XO_SF_OV_LB] + reserve_ov1;
    int min_shift = imin(shift1, shift2);
    int shift_over = shift2 - min_shift;
    reserve -= shift1 - min_shift;
    sf[XO_SF_OV_LB] = (i16)(sf[XO_SF_OV_LB] + (reserve_ov1 - shift_over));
    xo_adjust_scale_lp(m, 0, usb, 0, 6, reserve_ov1 - shift_over);
    xo_adjust_scale_lp(m, 0, usb, 6, 38, reserve);
    xo_adjust_scale_lp(lrows, 0, usb, 0, 2, reserve_ov1 - shift_over);
    for (int i = 0; i < 2; i++) memcpy(lpc + 128 * i, lrows + 64 * i, 32 * sizeof(i32));
    sf[XO_SF_LB] = (i16)(sf[XO_SF_LB] + reserve);
    save_lb_scale = sf[XO_SF_LB];
  }
  for (int l = 6; l < 38; l++) memset(m + 64 * l + 32, 0, 32 * sizeof(i32));
  if (apply) {
    i16 degree_alias[64];
    memset(degree_alias, 0, sizeof(degree_alias));
    i16 hf[XO_HF_PRM_WORDS];
    memcpy(hf, hfs, sizeof(hf));
    hf[XO_HF_START_IDX] = (i16)(border[0] * env[XO_ENV_TIME_STEP]);
    hf[XO_HF_STOP_IDX] = (i16)(env[XO_ENV_TIME_STEP] * ox_sub16_sat(border[num_env], env[XO_ENV_NUM_TIME_SLOTS]));
    for (int i = 0; i < 10; i++) hf[XO_HF_INVF_PREV + i] = misc[XO_SBR_MISC_INVF_PREV + i];
    hf[XO_HF_MAX_QMF_SUBBAND] = env[XO_ENV_MAX_QMF_SUBBAND];
    xo_low_pow_hf_generator(lpc, scratch, hf, bw_prev, degree_alias, max_samp_val);
    sf[XO_SF_HB] = (i16)(imin(sf[XO_SF_OV_LB], sf[XO_SF_LB]) - 2);
    i16 envp[XO_ENV_PRM_WORDS];
    memcpy(envp, env, sizeof(envp));
    envp[XO_ENV_MAX_QMF_SUBBAND_PREV] = misc[XO_SBR_MISC_MAX_QMF_PREV];
    int err = xo_calc_sbrenvelope_lp(env_rom, misc_rom, envp, sf, st + XO_SBR_ST_ENV, m, degree_alias);
    if (err) return err;
    for (int i = 0; i < hf[XO_HF_NUM_IF_BANDS]; i++) misc[XO_SBR_MISC_INVF_PREV + i] = hf[XO_HF_INVF + i];
    misc[XO_SBR_MISC_MAX_QMF_PREV] = env[XO_ENV_MAX_QMF_SUBBAND];
    misc[XO_SBR_MISC_END_POS_PREV] = border[num_env];
  } else {
    sf[XO_SF_HB] = (i16)save_lb_scale;
  }
  {
    int usb = misc[XO_SBR_MISC_CODEC_USB];
    for (int i = 0; i < 2; i++) memcpy(lpc + 128 * i, m + 64 * (30 + i), usb * sizeof(i32));
  }
  i32 ovsave[6 * 64];
  memcpy(ovsave, m + 32 * 64, sizeof(ovsave));
  i32 sfv[4] = {sf[XO_SF_OV_LB], sf[XO_SF_LB], sf[XO_SF_HB], sf[XO_SF_ST_SYN]};
  i32 off = st[XO_SBR_ST_SYN_POS], fpos = st[XO_SBR_ST_SYN_POS + 1];
  xo_synt_qmffilt_lp(qrom, m, st + XO_SBR_ST_SYN_STATES, &off, &fpos, sfv, misc[XO_SBR_MISC_SYN_LSB],
                     misc[XO_SBR_MISC_SYN_USB], 6, time_out, ch_out);
  st[XO_SBR_ST_SYN_POS] = (i16)off;
  st[XO_SBR_ST_SYN_POS + 1] = (i16)fpos;
  memcpy(ov, ovsave, sizeof(ovsave));
  sf[XO_SF_OV_LB] = (i16)save_lb_scale;
  return 0;
}
