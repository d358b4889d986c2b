/*
 * libxaac_b200/dropin/ixheaacd_b200_pack_ps_flt.h — reference-side half of the float parametric-stereo hand-over
 * (ixheaacd_esbr_apply_ps, decoder/ixheaacd_ps_dec_flt.c).  Compiled against the reference headers, like ixheaacd_b200_pack.h:
 *   b200_fps_pack_rom     ia_ps_tables_struct -> XAAC_FPSROM_* blob
 *   b200_fps_pack_state   ia_ps_dec_struct    -> XAAC_FPS_ST_* record     (b200_fps_unpack_state: the way back)
 *   b200_fps_side         the frame's PS parameters -> XAAC_FPS_SIDE_* record.  The mixing matrices of
 *                         ixheaacd_esbr_ps_apply_rotation (ps_dec_flt.c:920-1020) are evaluated HERE, with the C library's
 *                         double-precision cos / sin / atan2 the reference itself calls, so that the device only interpolates
 *                         and applies them; the smoothing history they advance is committed with b200_fps_commit once the
 *                         device call has succeeded.
 * Used by ixheaacd_b200_glue.c; the test shim oracle/ref_shim_fps.c includes it too.
 */
#ifndef IXHEAACD_B200_PACK_PS_FLT_H
#define IXHEAACD_B200_PACK_PS_FLT_H
#include <math.h>
#include <string.h>
#include "ixheaacd_b200_ref_headers.h"
#include "ixheaacd_hybrid.h"
#include "ixheaacd_ps_dec.h"
#include "xaac_b200.h"

static void b200_fps_pack_rom(float *rom, const ia_ps_tables_struct *t, const WORD16 *delay_sample_ser) {
  int32_t *irom = (int32_t *)rom;
  memset(rom, 0, XAAC_FPSROM_WORDS * 4);
  memcpy(rom + XAAC_FPSROM_P8, t->p8_13_20, 13 * 4);
  memcpy(rom + XAAC_FPSROM_P2, t->p2_13_20, 13 * 4);
  memcpy(rom + XAAC_FPSROM_COS2, t->cos_mod_2channel, 26 * 4);
  memcpy(rom + XAAC_FPSROM_CS8, t->cos_sin_mod_8channel, 208 * 4);
  memcpy(rom + XAAC_FPSROM_QF_RE, t->qmf_fract_delay_phase_factor_re, 64 * 4);
  memcpy(rom + XAAC_FPSROM_QF_IM, t->qmf_fract_delay_phase_factor_im, 64 * 4);
  memcpy(rom + XAAC_FPSROM_SUB_RE, t->frac_delay_phase_fac_qmf_sub_re_20, 12 * 4);
  memcpy(rom + XAAC_FPSROM_SUB_IM, t->frac_delay_phase_fac_qmf_sub_im_20, 12 * 4);
  memcpy(rom + XAAC_FPSROM_QSER_RE, t->qmf_ser_fract_delay_phase_factor_re, 192 * 4);
  memcpy(rom + XAAC_FPSROM_QSER_IM, t->qmf_ser_fract_delay_phase_factor_im, 192 * 4);
  memcpy(rom + XAAC_FPSROM_SSER_RE, t->frac_delay_phase_fac_ser_qmf_sub_re_20, 36 * 4);
  memcpy(rom + XAAC_FPSROM_SSER_IM, t->frac_delay_phase_fac_ser_qmf_sub_im_20, 36 * 4);
  memcpy(rom + XAAC_FPSROM_DECAY, t->all_pass_link_decay_ser, 3 * 4);
  for (int i = 0; i < 64; i++) irom[XAAC_FPSROM_QDELN + i] = t->qmf_delay_idx_tbl[i];
  for (int i = 0; i < 23; i++) irom[XAAC_FPSROM_GRB + i] = t->group_borders_20_tbl[i];
  for (int i = 0; i < 22; i++) irom[XAAC_FPSROM_BGM + i] = t->bin_group_map_20[i];
  for (int i = 0; i < 3; i++) irom[XAAC_FPSROM_DSER + i] = delay_sample_ser[i];
}

static void b200_fps_pack_state(float *st, const ia_ps_dec_struct *ps) {
  int32_t *ist = (int32_t *)st + XAAC_FPS_ST_IDX;
  memset(st, 0, XAAC_FPS_ST_WORDS * 4);
  /* both hybrids keep the same samples (each call shifts both, ps_dec_flt.c:432-437); bands 3, 4 only exist in the 34-band one */
  for (int b = 0; b < 5; b++) {
    memcpy(st + XAAC_FPS_ST_HYB + 12 * b, b < 3 ? ps->hyb_qmf_buf_re_20[b] : ps->hyb_qmf_buf_re_34[b], 48);
    memcpy(st + XAAC_FPS_ST_HYB + 60 + 12 * b, b < 3 ? ps->hyb_qmf_buf_im_20[b] : ps->hyb_qmf_buf_im_34[b], 48);
  }
  for (int r = 0; r < 2; r++) {
    memcpy(st + XAAC_FPS_ST_SUBDEL + 12 * r, ps->sub_qmf_delay_buf_re[r], 48);
    memcpy(st + XAAC_FPS_ST_SUBDEL + 24 + 12 * r, ps->sub_qmf_delay_buf_im[r], 48);
  }
  for (int m = 0; m < 3; m++)
    for (int k = 0; k < 5; k++) {
      memcpy(st + XAAC_FPS_ST_SERSUB + 12 * (5 * m + k), ps->ser_sub_qmf_dealy_buf_re[m][k], 48);
      memcpy(st + XAAC_FPS_ST_SERSUB + 180 + 12 * (5 * m + k), ps->ser_sub_qmf_dealy_buf_im[m][k], 48);
    }
  memcpy(st + XAAC_FPS_ST_QDEL, ps->qmf_delay_buf_re, 896 * 4);
  memcpy(st + XAAC_FPS_ST_QDEL + 896, ps->qmf_delay_buf_im, 896 * 4);
  memcpy(st + XAAC_FPS_ST_SERQ, ps->ser_qmf_delay_buf_re, 960 * 4);
  memcpy(st + XAAC_FPS_ST_SERQ + 960, ps->ser_qmf_delay_buf_im, 960 * 4);
  memcpy(st + XAAC_FPS_ST_BINS, ps->peak_decay_fast_bin, 80);
  memcpy(st + XAAC_FPS_ST_BINS + 20, ps->prev_nrg_bin, 80);
  memcpy(st + XAAC_FPS_ST_BINS + 40, ps->prev_peak_diff_bin, 80);
  ist[0] = ps->delay_buf_idx;
  for (int m = 0; m < 3; m++) ist[1 + m] = ps->delay_buf_idx_ser[m];
  for (int i = 0; i < 64; i++) ist[4 + i] = ps->delay_qmf_delay_buf_idx[i];
}

static void b200_fps_unpack_state(const float *st, ia_ps_dec_struct *ps) {
  const int32_t *ist = (const int32_t *)st + XAAC_FPS_ST_IDX;
  for (int b = 0; b < 5; b++) {
    if (b < 3) {
      memcpy(ps->hyb_qmf_buf_re_20[b], st + XAAC_FPS_ST_HYB + 12 * b, 48);
      memcpy(ps->hyb_qmf_buf_im_20[b], st + XAAC_FPS_ST_HYB + 60 + 12 * b, 48);
    }
    memcpy(ps->hyb_qmf_buf_re_34[b], st + XAAC_FPS_ST_HYB + 12 * b, 48);
    memcpy(ps->hyb_qmf_buf_im_34[b], st + XAAC_FPS_ST_HYB + 60 + 12 * b, 48);
  }
  for (int r = 0; r < 2; r++) {
    memcpy(ps->sub_qmf_delay_buf_re[r], st + XAAC_FPS_ST_SUBDEL + 12 * r, 48);
    memcpy(ps->sub_qmf_delay_buf_im[r], st + XAAC_FPS_ST_SUBDEL + 24 + 12 * r, 48);
  }
  for (int m = 0; m < 3; m++)
    for (int k = 0; k < 5; k++) {
      memcpy(ps->ser_sub_qmf_dealy_buf_re[m][k], st + XAAC_FPS_ST_SERSUB + 12 * (5 * m + k), 48);
      memcpy(ps->ser_sub_qmf_dealy_buf_im[m][k], st + XAAC_FPS_ST_SERSUB + 180 + 12 * (5 * m + k), 48);
    }
  memcpy(ps->qmf_delay_buf_re, st + XAAC_FPS_ST_QDEL, 896 * 4);
  memcpy(ps->qmf_delay_buf_im, st + XAAC_FPS_ST_QDEL + 896, 896 * 4);
  memcpy(ps->ser_qmf_delay_buf_re, st + XAAC_FPS_ST_SERQ, 960 * 4);
  memcpy(ps->ser_qmf_delay_buf_im, st + XAAC_FPS_ST_SERQ + 960, 960 * 4);
  memcpy(ps->peak_decay_fast_bin, st + XAAC_FPS_ST_BINS, 80);
  memcpy(ps->prev_nrg_bin, st + XAAC_FPS_ST_BINS + 20, 80);
  memcpy(ps->prev_peak_diff_bin, st + XAAC_FPS_ST_BINS + 40, 80);
  ps->delay_buf_idx = (WORD16)ist[0];
  for (int m = 0; m < 3; m++) ps->delay_buf_idx_ser[m] = (WORD16)ist[1 + m];
  for (int i = 0; i < 64; i++) ps->delay_qmf_delay_buf_idx[i] = ist[4 + i];
}

typedef struct {
  float h_last[8][20];                       /* the h*_vec of the last envelope = the next frame's h*_prev */
  WORD32 ipd1[17], opd1[17], ipd2[17], opd2[17];
} b200_fps_commit_rec;

/* Returns 0, or -1 when the frame is outside what the device path covers (34 stereo bands, PCA rotation, the ps_mode debug
 * switches, borders that do not tile the 32 slots): the caller then leaves the frame to the reference. */
static int b200_fps_side(float *side, b200_fps_commit_rec *cm, const ia_ps_dec_struct *ps, const ia_ps_tables_struct *t, int usb) {
  int32_t *iside = (int32_t *)side;
  const int ne = ps->num_env;
  if (ps->use_34_st_bands || ps->use_34_st_bands_prev || ps->use_pca_rot_flg || ps->ps_mode || ps->num_sub_samples != 32 ||
      ps->num_chans != 64 || ne < 1 || ne > 5 || ps->freq_res_ipd < 0 || ps->freq_res_ipd > 2)
    return -1;
  if (ps->border_position[0] != 0 || ps->border_position[ne] != 32) return -1;
  for (int e = 0; e < ne; e++)
    if (ps->border_position[e + 1] <= ps->border_position[e]) return -1;
  memset(side, 0, XAAC_FPS_SIDE_WORDS * 4);
  iside[XAAC_FPS_SIDE_NUM_ENV] = ne;
  for (int e = 0; e <= ne; e++) iside[XAAC_FPS_SIDE_BORDER + e] = ps->border_position[e];
  iside[XAAC_FPS_SIDE_USB] = usb;
  const float *prev[8] = {ps->h11_re_prev, ps->h12_re_prev, ps->h21_re_prev, ps->h22_re_prev,
                          ps->h11_im_prev, ps->h12_im_prev, ps->h21_im_prev, ps->h22_im_prev};
  for (int c = 0; c < 8; c++) memcpy(side + XAAC_FPS_SIDE_H + 20 * c, prev[c], 80);
  memcpy(cm->ipd1, ps->ipd_idx_map_1, sizeof(cm->ipd1));
  memcpy(cm->opd1, ps->opd_idx_map_1, sizeof(cm->opd1));
  memcpy(cm->ipd2, ps->ipd_idx_map_2, sizeof(cm->ipd2));
  memcpy(cm->opd2, ps->opd_idx_map_2, sizeof(cm->opd2));
  const int steps = ps->iid_quant ? NUM_IID_STEPS_FINE : NUM_IID_STEPS;
  const FLOAT32 *sf = ps->iid_quant ? t->scale_factors_fine_flt : t->scale_factors_flt;
  const int ipd_bins = t->ipd_bins_tbl[ps->freq_res_ipd];
  if (ipd_bins > 17) return -1;
  const FLOAT32 ang = IPD_SCALE_FACTOR * 2.0f; /* == OPD_SCALE_FACTOR * 2.0f */
  for (int e = 0; e < ne; e++) {
    float *h = side + XAAC_FPS_SIDE_H + 160 * (e + 1);
    for (int b = 0; b < 20; b++) {
      const int iid = ps->iid_par_table[e][b], icc = ps->icc_par_table[e][b];
      if (iid < -steps || iid > steps || icc < 0 || icc >= NUM_ICC_LEVELS) return -1;
      const FLOAT32 sr = sf[steps + iid], sl = sf[steps - iid];
      const FLOAT32 alpha = t->alphas[icc];
      const FLOAT32 beta = alpha * (sr - sl) / PSC_SQRT2F;
      FLOAT32 r11 = (FLOAT32)(sl * cos(beta + alpha));
      FLOAT32 r12 = (FLOAT32)(sr * cos(beta - alpha));
      FLOAT32 r21 = (FLOAT32)(sl * sin(beta + alpha));
      FLOAT32 r22 = (FLOAT32)(sr * sin(beta - alpha));
      FLOAT32 i11 = 0.0f, i12 = 0.0f, i21 = 0.0f, i22 = 0.0f;
      if (b < ipd_bins) {
        /* the phases are smoothed over the two previous envelopes before they rotate the matrix */
        FLOAT32 ipd = ang * ps->ipd_idx_map[e][b], opd = ang * ps->opd_idx_map[e][b];
        const FLOAT32 ipd_1 = ang * cm->ipd1[b], opd_1 = ang * cm->opd1[b];
        const FLOAT32 ipd_2 = ang * cm->ipd2[b], opd_2 = ang * cm->opd2[b];
        FLOAT32 lre = (FLOAT32)cos(ipd), lim = (FLOAT32)sin(ipd), rre = (FLOAT32)cos(opd), rim = (FLOAT32)sin(opd);
        lre += PHASE_SMOOTH_HIST1 * (FLOAT32)cos(ipd_1);
        lim += PHASE_SMOOTH_HIST1 * (FLOAT32)sin(ipd_1);
        rre += PHASE_SMOOTH_HIST1 * (FLOAT32)cos(opd_1);
        rim += PHASE_SMOOTH_HIST1 * (FLOAT32)sin(opd_1);
        lre += PHASE_SMOOTH_HIST2 * (FLOAT32)cos(ipd_2);
        lim += PHASE_SMOOTH_HIST2 * (FLOAT32)sin(ipd_2);
        rre += PHASE_SMOOTH_HIST2 * (FLOAT32)cos(opd_2);
        rim += PHASE_SMOOTH_HIST2 * (FLOAT32)sin(opd_2);
        ipd = (FLOAT32)atan2(lim, lre);
        opd = (FLOAT32)atan2(rim, rre);
        lre = (FLOAT32)cos(opd);
        lim = (FLOAT32)sin(opd);
        opd -= ipd;
        rre = (FLOAT32)cos(opd);
        rim = (FLOAT32)sin(opd);
        i11 = r11 * lim; i12 = r12 * rim; i21 = r21 * lim; i22 = r22 * rim;
        r11 *= lre; r12 *= rre; r21 *= lre; r22 *= rre;
      }
      h[b] = r11; h[20 + b] = r12; h[40 + b] = r21; h[60 + b] = r22;
      h[80 + b] = i11; h[100 + b] = i12; h[120 + b] = i21; h[140 + b] = i22;
    }
    for (int b = 0; b < ipd_bins; b++) {
      cm->ipd2[b] = cm->ipd1[b]; cm->opd2[b] = cm->opd1[b];
      cm->ipd1[b] = ps->ipd_idx_map[e][b]; cm->opd1[b] = ps->opd_idx_map[e][b];
    }
  }
  memcpy(cm->h_last, side + XAAC_FPS_SIDE_H + 160 * ne, sizeof(cm->h_last));
  return 0;
}

static void b200_fps_commit(ia_ps_dec_struct *ps, const b200_fps_commit_rec *cm, ia_ps_tables_struct *t) {
  float *prev[8] = {ps->h11_re_prev, ps->h12_re_prev, ps->h21_re_prev, ps->h22_re_prev,
                    ps->h11_im_prev, ps->h12_im_prev, ps->h21_im_prev, ps->h22_im_prev};
  float *vec[8] = {ps->h11_re_vec, ps->h12_re_vec, ps->h21_re_vec, ps->h22_re_vec,
                   ps->h11_im_vec, ps->h12_im_vec, ps->h21_im_vec, ps->h22_im_vec};
  for (int c = 0; c < 8; c++) {
    memcpy(prev[c], cm->h_last[c], 80);
    memcpy(vec[c], cm->h_last[c], 80);
  }
  memcpy(ps->ipd_idx_map_1, cm->ipd1, sizeof(cm->ipd1));
  memcpy(ps->opd_idx_map_1, cm->opd1, sizeof(cm->opd1));
  memcpy(ps->ipd_idx_map_2, cm->ipd2, sizeof(cm->ipd2));
  memcpy(ps->opd_idx_map_2, cm->opd2, sizeof(cm->opd2));
  /* what ixheaacd_esbr_apply_ps leaves in the instance besides the signal state (ps_dec_flt.c:408-416, 503) */
  ps->ptr_group_borders = (WORD32 *)&t->group_borders_20_tbl[0];
  ps->ptr_bins_group_map = (WORD32 *)&t->bin_group_map_20[0];
  ps->ptr_hybrid = &ps->str_flt_hybrid20;
  ps->num_groups = NUM_IID_GROUPS;
  ps->num_sub_qmf_groups = SUBQMF_GROUPS;
  ps->num_bins = NUM_MID_RES_BINS;
  ps->first_delay_gr = SUBQMF_GROUPS;
  ps->use_34_st_bands_prev = ps->use_34_st_bands;
}
#endif
