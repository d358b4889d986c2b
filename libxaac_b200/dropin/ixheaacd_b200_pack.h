/*
 * libxaac_b200/dropin/ixheaacd_b200_pack.h — reference-side half of the drop-in (compiled inside / against the reference tree).
 * Conversions between the reference's pointer-rich SBR / PS structs and the flat WORD16 records of the C-ABI
 * (include/xaac_b200.h, XAAC_SIDE_* / XAAC_SBR_ST_* / XAAC_PS_ST_* offsets): struct -> record before a stage call, record ->
 * struct afterwards.  Used by the link-time stage overrides in ixheaacd_b200_glue.c; the test infrastructure under oracle/
 * (taps, the shim that drives the compiled reference from records) includes this file too.
 */
#ifndef IXHEAACD_B200_PACK_H
#define IXHEAACD_B200_PACK_H
#include "ixheaacd_b200_ref_headers.h"
#include "ixheaacd_env_calc.h"
#include "ixheaac_sbr_const.h"
#include "ixheaacd_pvc_dec.h"
#include "ixheaacd_sbr_dec.h"
#include "xaac_b200.h"

static void pack_env_prm(int16_t *prm, const ia_sbr_header_data_struct *h, const ia_sbr_frame_info_data_struct *f,
                         const ia_sbr_prev_frame_data_struct *pv) {
  const ia_freq_band_data_struct *fb = h->pstr_freq_band_data;
  const ia_frame_info_struct *fi = &f->str_frame_info_details;
  memset(prm, 0, XAAC_ENV_PRM_WORDS * sizeof(int16_t));
  prm[XAAC_ENV_NUM_TIME_SLOTS] = h->num_time_slots;
  prm[XAAC_ENV_TIME_STEP] = h->time_step;
  prm[XAAC_ENV_CHANNEL_MODE] = (int16_t)h->channel_mode;
  prm[XAAC_ENV_LIMITER_GAINS] = h->limiter_gains;
  prm[XAAC_ENV_INTERPOL_FREQ] = h->interpol_freq;
  prm[XAAC_ENV_SMOOTHING_MODE] = h->smoothing_mode;
  prm[XAAC_ENV_NUM_SF_LO] = fb->num_sf_bands[0];
  prm[XAAC_ENV_NUM_SF_HI] = fb->num_sf_bands[1];
  prm[XAAC_ENV_NUM_NF_BANDS] = fb->num_nf_bands;
  prm[XAAC_ENV_SUB_BAND_START] = fb->sub_band_start;
  prm[XAAC_ENV_SUB_BAND_END] = fb->sub_band_end;
  prm[XAAC_ENV_NUM_LF_BANDS] = fb->num_lf_bands;
  prm[XAAC_ENV_NUM_ENV] = fi->num_env;
  prm[XAAC_ENV_TRANSIENT_ENV] = fi->transient_env;
  prm[XAAC_ENV_MAX_QMF_SUBBAND] = (int16_t)f->max_qmf_subband_aac;
  prm[XAAC_ENV_MAX_QMF_SUBBAND_PREV] = (int16_t)pv->max_qmf_subband_aac;
  for (int i = 0; i < MAX_ENVELOPES + 1; i++) prm[XAAC_ENV_BORDER_VEC + i] = fi->border_vec[i];
  for (int i = 0; i < MAX_ENVELOPES; i++) prm[XAAC_ENV_FREQ_RES + i] = fi->freq_res[i];
  for (int i = 0; i < MAX_NOISE_ENVELOPES + 1; i++) prm[XAAC_ENV_NOISE_BORDER_VEC + i] = fi->noise_border_vec[i];
  for (int i = 0; i < MAX_NUM_LIMITERS + 1; i++) prm[XAAC_ENV_LIM_TBL + i] = fb->freq_band_tbl_lim[i];
  for (int i = 0; i < MAX_FREQ_COEFFS / 2 + 1; i++) prm[XAAC_ENV_FREQ_LO + i] = fb->freq_band_tbl_lo[i];
  for (int i = 0; i < MAX_FREQ_COEFFS + 1; i++) prm[XAAC_ENV_FREQ_HI + i] = fb->freq_band_tbl_hi[i];
  for (int i = 0; i < MAX_NOISE_COEFFS + 1; i++) prm[XAAC_ENV_FREQ_NOISE + i] = fb->freq_band_tbl_noise[i];
  for (int i = 0; i < MAX_NUM_NOISE_VALUES; i++) prm[XAAC_ENV_NOISE_FLOOR + i] = f->int_noise_floor[i];
  for (int i = 0; i < MAX_FREQ_COEFFS; i++) prm[XAAC_ENV_ADD_HARMONICS + i] = (int16_t)f->add_harmonics[i];
  for (int i = 0; i < MAX_NUM_ENVELOPE_VALUES; i++) prm[XAAC_ENV_SF_ARR + i] = f->int_env_sf_arr[i];
}

static void unpack_env_prm(const int16_t *prm, ia_sbr_header_data_struct *h, ia_freq_band_data_struct *fb,
                           ia_sbr_frame_info_data_struct *f, ia_sbr_prev_frame_data_struct *pv) {
  ia_frame_info_struct *fi = &f->str_frame_info_details;
  h->pstr_freq_band_data = fb;
  h->num_time_slots = prm[XAAC_ENV_NUM_TIME_SLOTS];
  h->time_step = prm[XAAC_ENV_TIME_STEP];
  h->channel_mode = prm[XAAC_ENV_CHANNEL_MODE];
  h->limiter_gains = prm[XAAC_ENV_LIMITER_GAINS];
  h->interpol_freq = prm[XAAC_ENV_INTERPOL_FREQ];
  h->smoothing_mode = prm[XAAC_ENV_SMOOTHING_MODE];
  fb->num_sf_bands[0] = prm[XAAC_ENV_NUM_SF_LO];
  fb->num_sf_bands[1] = prm[XAAC_ENV_NUM_SF_HI];
  fb->num_nf_bands = prm[XAAC_ENV_NUM_NF_BANDS];
  fb->sub_band_start = prm[XAAC_ENV_SUB_BAND_START];
  fb->sub_band_end = prm[XAAC_ENV_SUB_BAND_END];
  fb->num_lf_bands = prm[XAAC_ENV_NUM_LF_BANDS];
  fb->freq_band_table[0] = fb->freq_band_tbl_lo;
  fb->freq_band_table[1] = fb->freq_band_tbl_hi;
  fi->num_env = prm[XAAC_ENV_NUM_ENV];
  fi->transient_env = prm[XAAC_ENV_TRANSIENT_ENV];
  f->max_qmf_subband_aac = prm[XAAC_ENV_MAX_QMF_SUBBAND];
  pv->max_qmf_subband_aac = prm[XAAC_ENV_MAX_QMF_SUBBAND_PREV];
  for (int i = 0; i < MAX_ENVELOPES + 1; i++) fi->border_vec[i] = prm[XAAC_ENV_BORDER_VEC + i];
  for (int i = 0; i < MAX_ENVELOPES; i++) fi->freq_res[i] = prm[XAAC_ENV_FREQ_RES + i];
  for (int i = 0; i < MAX_NOISE_ENVELOPES + 1; i++) fi->noise_border_vec[i] = prm[XAAC_ENV_NOISE_BORDER_VEC + i];
  for (int i = 0; i < MAX_NUM_LIMITERS + 1; i++) fb->freq_band_tbl_lim[i] = prm[XAAC_ENV_LIM_TBL + i];
  for (int i = 0; i < MAX_FREQ_COEFFS / 2 + 1; i++) fb->freq_band_tbl_lo[i] = prm[XAAC_ENV_FREQ_LO + i];
  for (int i = 0; i < MAX_FREQ_COEFFS + 1; i++) fb->freq_band_tbl_hi[i] = prm[XAAC_ENV_FREQ_HI + i];
  for (int i = 0; i < MAX_NOISE_COEFFS + 1; i++) fb->freq_band_tbl_noise[i] = prm[XAAC_ENV_FREQ_NOISE + i];
  for (int i = 0; i < MAX_NUM_NOISE_VALUES; i++) f->int_noise_floor[i] = prm[XAAC_ENV_NOISE_FLOOR + i];
  for (int i = 0; i < MAX_FREQ_COEFFS; i++) f->add_harmonics[i] = prm[XAAC_ENV_ADD_HARMONICS + i];
  for (int i = 0; i < MAX_NUM_ENVELOPE_VALUES; i++) f->int_env_sf_arr[i] = prm[XAAC_ENV_SF_ARR + i];
}

static void pack_sf(int16_t *sf, const ia_sbr_scale_fact_struct *s) {
  sf[0] = s->lb_scale; sf[1] = s->st_lb_scale; sf[2] = s->ov_lb_scale; sf[3] = s->hb_scale;
  sf[4] = s->ov_hb_scale; sf[5] = s->st_syn_scale; sf[6] = s->ps_scale; sf[7] = 0;
}
static void unpack_sf(const int16_t *sf, ia_sbr_scale_fact_struct *s) {
  s->lb_scale = sf[0]; s->st_lb_scale = sf[1]; s->ov_lb_scale = sf[2]; s->hb_scale = sf[3];
  s->ov_hb_scale = sf[4]; s->st_syn_scale = sf[5]; s->ps_scale = sf[6];
}

static void pack_env_state(int16_t *st, const ia_sbr_calc_env_struct *e) {
  memset(st, 0, XAAC_ENV_ST_WORDS * sizeof(int16_t));
  memcpy(st + XAAC_ENV_ST_FILT_ME, e->filt_buf_me, 2 * MAX_FREQ_COEFFS * sizeof(int16_t));
  memcpy(st + XAAC_ENV_ST_FILT_NOISE, e->filt_buf_noise_m, MAX_FREQ_COEFFS * sizeof(int16_t));
  st[XAAC_ENV_ST_NOISE_E] = (int16_t)e->filt_buf_noise_e;
  st[XAAC_ENV_ST_START_UP] = (int16_t)e->start_up;
  st[XAAC_ENV_ST_PH_INDEX] = e->ph_index;
  st[XAAC_ENV_ST_TRANS_PREV] = e->tansient_env_prev;
  st[XAAC_ENV_ST_HARM_INDEX] = e->harm_index;
  for (int i = 0; i < MAX_FREQ_COEFFS; i++) st[XAAC_ENV_ST_HARM_PREV + i] = e->harm_flags_prev[i];
}
/* e->filt_buf_me / filt_buf_noise_m must already point at caller-owned buffers */
static void unpack_env_state(const int16_t *st, ia_sbr_calc_env_struct *e) {
  memcpy(e->filt_buf_me, st + XAAC_ENV_ST_FILT_ME, 2 * MAX_FREQ_COEFFS * sizeof(int16_t));
  memcpy(e->filt_buf_noise_m, st + XAAC_ENV_ST_FILT_NOISE, MAX_FREQ_COEFFS * sizeof(int16_t));
  e->filt_buf_noise_e = st[XAAC_ENV_ST_NOISE_E];
  e->start_up = st[XAAC_ENV_ST_START_UP];
  e->ph_index = st[XAAC_ENV_ST_PH_INDEX];
  e->tansient_env_prev = st[XAAC_ENV_ST_TRANS_PREV];
  e->harm_index = st[XAAC_ENV_ST_HARM_INDEX];
  for (int i = 0; i < MAX_FREQ_COEFFS; i++) e->harm_flags_prev[i] = (WORD8)st[XAAC_ENV_ST_HARM_PREV + i];
}

/* ---- whole-stage records (XAAC_SBR_ST_*, XAAC_PS_ST_*, XAAC_SIDE_*) ------------------------------------------------ */
static void pack_hf_settings(int16_t *prm, const ia_transposer_settings_struct *set) {
  prm[XAAC_HF_NUM_PATCHES] = set->num_patches; prm[XAAC_HF_START_PATCH] = set->start_patch;
  prm[XAAC_HF_STOP_PATCH] = set->stop_patch; prm[XAAC_HF_NUM_COLUMNS] = set->num_columns;
  for (int i = 0; i < MAX_NUM_NOISE_VALUES; i++) prm[XAAC_HF_BW_BORDERS + i] = set->bw_borders[i];
  for (int p = 0; p < MAX_NUM_PATCHES; p++) {
    int16_t *q = prm + XAAC_HF_PATCH + 6 * p;
    q[0] = set->str_patch_param[p].src_start_band; q[1] = set->str_patch_param[p].src_end_band;
    q[2] = set->str_patch_param[p].guard_start_band; q[3] = set->str_patch_param[p].dst_start_band;
    q[4] = set->str_patch_param[p].dst_end_band; q[5] = set->str_patch_param[p].num_bands_in_patch;
  }
}
static void unpack_hf_settings(const int16_t *prm, ia_transposer_settings_struct *set) {
  set->num_patches = prm[XAAC_HF_NUM_PATCHES]; set->start_patch = prm[XAAC_HF_START_PATCH];
  set->stop_patch = prm[XAAC_HF_STOP_PATCH]; set->num_columns = prm[XAAC_HF_NUM_COLUMNS];
  for (int i = 0; i < MAX_NUM_NOISE_VALUES; i++) set->bw_borders[i] = prm[XAAC_HF_BW_BORDERS + i];
  for (int p = 0; p < MAX_NUM_PATCHES; p++) {
    const int16_t *q = prm + XAAC_HF_PATCH + 6 * p;
    set->str_patch_param[p].src_start_band = q[0]; set->str_patch_param[p].src_end_band = q[1];
    set->str_patch_param[p].guard_start_band = q[2]; set->str_patch_param[p].dst_start_band = q[3];
    set->str_patch_param[p].dst_end_band = q[4]; set->str_patch_param[p].num_bands_in_patch = q[5];
  }
}

static void pack_side(int16_t *side, const ia_sbr_dec_struct *d, const ia_sbr_header_data_struct *h,
                      const ia_sbr_frame_info_data_struct *f, const ia_sbr_prev_frame_data_struct *pv,
                      const ia_ps_dec_struct *ps, int apply) {
  memset(side, 0, XAAC_SIDE_WORDS * sizeof(int16_t));
  pack_env_prm(side + XAAC_SIDE_ENV, h, f, pv);
  int16_t *hf = side + XAAC_SIDE_HF;
  pack_hf_settings(hf, d->str_hf_generator.pstr_settings);
  hf[XAAC_HF_FACTOR] = h->time_step;
  hf[XAAC_HF_NUM_IF_BANDS] = h->pstr_freq_band_data->num_if_bands;
  for (int i = 0; i < MAX_NUM_NOISE_VALUES; i++) hf[XAAC_HF_INVF + i] = (int16_t)f->sbr_invf_mode[i];
  side[XAAC_SIDE_APPLY] = (int16_t)apply;
  side[XAAC_SIDE_PS] = (int16_t)(ps != NULL && h->channel_mode == PS_STEREO);
  if (ps) {
    int16_t *pp = side + XAAC_SIDE_PS_PRM;
    pp[XAAC_PS_PRM_IID_QUANT] = (int16_t)ps->iid_quant;
    pp[XAAC_PS_PRM_NUM_ENV] = ps->num_env;
    for (int i = 0; i < MAXIM_NUM_OF_PS_ENVLOPS + 2; i++) pp[XAAC_PS_PRM_BORDER + i] = ps->border_position[i];
    memcpy(pp + XAAC_PS_PRM_IID, ps->iid_par_table, 238 * sizeof(int16_t));
    memcpy(pp + XAAC_PS_PRM_ICC, ps->icc_par_table, 238 * sizeof(int16_t));
  }
}
static void unpack_side_ps(const int16_t *side, ia_ps_dec_struct *ps) {
  const int16_t *pp = side + XAAC_SIDE_PS_PRM;
  ps->iid_quant = pp[XAAC_PS_PRM_IID_QUANT];
  ps->num_env = pp[XAAC_PS_PRM_NUM_ENV];
  for (int i = 0; i < MAXIM_NUM_OF_PS_ENVLOPS + 2; i++) ps->border_position[i] = pp[XAAC_PS_PRM_BORDER + i];
  memcpy(ps->iid_par_table, pp + XAAC_PS_PRM_IID, 238 * sizeof(int16_t));
  memcpy(ps->icc_par_table, pp + XAAC_PS_PRM_ICC, 238 * sizeof(int16_t));
}

static void pack_sbr_state_lp(int16_t *st, const ia_sbr_dec_struct *d, const ia_sbr_prev_frame_data_struct *pv,
                              int low_pow) {
  const ia_sbr_qmf_filter_bank_struct *a = &d->str_codec_qmf_bank, *s = &d->str_synthesis_qmf_bank;
  memset(st, 0, XAAC_SBR_ST_WORDS * sizeof(int16_t));
  memcpy(st + XAAC_SBR_ST_ANAL_STATES, a->anal_filter_states, 320 * sizeof(int16_t));
  st[XAAC_SBR_ST_ANAL_POS] = (int16_t)(a->core_samples_buffer - a->anal_filter_states);
  st[XAAC_SBR_ST_ANAL_POS + 1] = (int16_t)(a->filter_pos - a->analy_win_coeff);
  st[XAAC_SBR_ST_SYN_POS] = s->ixheaacd_drc_offset;
  st[XAAC_SBR_ST_SYN_POS + 1] = (int16_t)(s->filter_pos_syn - s->p_filter);
  pack_sf(st + XAAC_SBR_ST_SF, &d->str_sbr_scale_fact);
  int16_t *misc = st + XAAC_SBR_ST_MISC;
  misc[XAAC_SBR_MISC_MAX_QMF_PREV] = (int16_t)pv->max_qmf_subband_aac;
  misc[XAAC_SBR_MISC_END_POS_PREV] = pv->end_position;
  for (int i = 0; i < MAX_NUM_NOISE_VALUES; i++) misc[XAAC_SBR_MISC_INVF_PREV + i] = (int16_t)pv->sbr_invf_mode[i];
  misc[XAAC_SBR_MISC_CODEC_USB] = a->usb;
  misc[XAAC_SBR_MISC_SYN_LSB] = s->lsb;
  misc[XAAC_SBR_MISC_SYN_USB] = s->usb;
  pack_env_state(st + XAAC_SBR_ST_ENV, &d->str_sbr_calc_env);
  memcpy(st + XAAC_SBR_ST_SYN_STATES, s->filter_states, 1280 * sizeof(int16_t));
  memcpy(st + XAAC_SBR_ST_BW_PREV, d->str_hf_generator.bw_array_prev, 6 * sizeof(int32_t));
  int32_t *lpc = (int32_t *)(st + XAAC_SBR_ST_LPC);
  for (int i = 0; i < 2; i++) {
    /* the reference allocates NO_ANALYSIS_CHANNELS (32) words per LPC state row (sbrdec_initfuncs.c:946-968) */
    memcpy(lpc + 128 * i, d->str_hf_generator.lpc_filt_states_real[i], 32 * sizeof(int32_t));
    if (!low_pow && d->str_hf_generator.lpc_filt_states_imag[i])
      memcpy(lpc + 128 * i + 64, d->str_hf_generator.lpc_filt_states_imag[i], 32 * sizeof(int32_t));
  }
  /* low power: 6 real slots of 64 words (raw, in the reference's own order); HQ: 6 x (re[64] | im[64]) */
  memcpy(st + XAAC_SBR_ST_OV, d->ptr_sbr_overlap_buf, (low_pow ? 6 * 64 : 6 * 128) * sizeof(int32_t));
}
static void pack_sbr_state(int16_t *st, const ia_sbr_dec_struct *d, const ia_sbr_prev_frame_data_struct *pv) {
  pack_sbr_state_lp(st, d, pv, 0);
}

static void pack_ps_state(int16_t *p, const ia_ps_dec_struct *ps, const ia_sbr_qmf_filter_bank_struct *bank_r,
                          const ia_sbr_scale_fact_struct *sf_r) {
  memset(p, 0, XAAC_PS_ST_WORDS * sizeof(int16_t));
  memcpy(p + XAAC_PS_ST_AP, ps->delay_buf_qmf_ap_re_im, 128 * sizeof(int16_t));
  memcpy(p + XAAC_PS_ST_LD, ps->delay_buf_qmf_ld_re_im, 336 * sizeof(int16_t));
  memcpy(p + XAAC_PS_ST_SD, ps->delay_buf_qmf_sd_re_im, 58 * sizeof(int16_t));
  memcpy(p + XAAC_PS_ST_SER, ps->delay_buf_qmf_ser_re_im, 960 * sizeof(int16_t));
  memcpy(p + XAAC_PS_ST_SUB, ps->delay_buf_qmf_sub_re_im, 64 * sizeof(int16_t));
  memcpy(p + XAAC_PS_ST_SUB_SER, ps->delay_buf_qmf_sub_ser_re_im, 480 * sizeof(int16_t));
  int16_t *hv = p + XAAC_PS_ST_HVEC;
  memcpy(hv, ps->h11_h12_vec, 96); memcpy(hv + 48, ps->h21_h22_vec, 96); memcpy(hv + 96, ps->H11_H12, 96);
  memcpy(hv + 144, ps->H21_H22, 96); memcpy(hv + 192, ps->delta_h11_h12, 96); memcpy(hv + 240, ps->delta_h21_h22, 96);
  int16_t *idx = p + XAAC_PS_ST_IDX;
  for (int i = 0; i < 3; i++) idx[XAAC_PS_IDX_SER + i] = ps->delay_buf_idx_ser[i];
  idx[XAAC_PS_IDX_DELAY] = ps->delay_buf_idx; idx[XAAC_PS_IDX_DELAY_LONG] = ps->delay_buf_idx_long;
  idx[XAAC_PS_IDX_SCALE] = ps->delay_buffer_scale; idx[XAAC_PS_IDX_USB] = ps->usb;
  idx[XAAC_PS_IDX_LSB_R] = bank_r->lsb; idx[XAAC_PS_IDX_USB_R] = bank_r->usb;
  int32_t *pk = (int32_t *)(p + XAAC_PS_ST_PEAK);
  memcpy(pk, ps->peak_decay_diff, 80); memcpy(pk + 20, ps->energy_prev, 80); memcpy(pk + 40, ps->peak_decay_diff_prev, 80);
  int32_t *hy = (int32_t *)(p + XAAC_PS_ST_HYB);
  for (int b = 0; b < 3; b++) {
    memcpy(hy + 24 * b, ps->str_hybrid.ptr_qmf_buf_re[b], 48);
    memcpy(hy + 24 * b + 12, ps->str_hybrid.ptr_qmf_buf_im[b], 48);
  }
  memcpy(p + XAAC_PS_ST_SYN_STATES_R, bank_r->filter_states, 1280 * sizeof(int16_t));
  p[XAAC_PS_ST_SYN_POS_R] = bank_r->ixheaacd_drc_offset;
  p[XAAC_PS_ST_SYN_POS_R + 1] = (int16_t)(bank_r->filter_pos_syn - bank_r->p_filter);
  pack_sf(p + XAAC_PS_ST_SF_R, sf_r);
}

/* Everything ixheaacd_sbr_dec needs, rebuilt from the flat records (used by the shim that drives the compiled
 * reference on arbitrary records).  All pointers point into this object. */
typedef struct {
  ia_sbr_dec_struct d;
  ia_sbr_header_data_struct h;
  ia_freq_band_data_struct fb;
  struct { ia_sbr_frame_info_data_struct f; WORD16 extra[512]; } __attribute__((aligned(8))) fd;
  ia_sbr_prev_frame_data_struct pv;
  ia_transposer_settings_struct set;
  ia_ps_dec_struct ps;
  ia_sbr_qmf_filter_bank_struct bank_r;
  ia_sbr_scale_fact_struct sf_r;
  ia_sbr_tables_struct tabs;
  WORD16 anal_states[320], syn_states[1280], syn_states_r[1280];
  WORD16 filt_me[2 * MAX_FREQ_COEFFS], filt_noise[MAX_FREQ_COEFFS];
  WORD32 lpc_r[2][64], lpc_i[2][64], ov[6 * 128];
  WORD32 lpc_lp[2][32]; /* low power: the reference walks from row 0 to row 1 with a stride of 32 words */
  WORD16 ps_ap[2][64], ps_ld[14][24], ps_sd[64], ps_ser[5][3][64];
  WORD32 ps_hyb_io[64], ps_work[32], ps_temp[16], ps_qbuf[3][2][12], ps_peak[60];
  WORD32 qmf_out[32][128];
  WORD32 work[64 * 128];
} ref_sbr_ctx;

static void unpack_sbr_ctx_lp(ref_sbr_ctx *c, const int16_t *side, const int16_t *st, const int16_t *p, int low_pow);
static void unpack_sbr_ctx(ref_sbr_ctx *c, const int16_t *side, const int16_t *st, const int16_t *p) {
  unpack_sbr_ctx_lp(c, side, st, p, 0);
}
static void unpack_sbr_ctx_lp(ref_sbr_ctx *c, const int16_t *side, const int16_t *st, const int16_t *p, int low_pow) {
  memset(c, 0, sizeof(*c));
  ia_qmf_dec_tables_struct *qt = (ia_qmf_dec_tables_struct *)&ixheaacd_aac_qmf_dec_tables;
  c->tabs.env_calc_tables_ptr = (ia_env_calc_tables_struct *)&ixheaacd_aac_dec_env_calc_tables;
  c->tabs.qmf_dec_tables_ptr = qt;
  c->tabs.ps_tables_ptr = (ia_ps_tables_struct *)&ixheaacd_aac_dec_ps_tables;
  c->tabs.sbr_rand_ph = c->tabs.env_calc_tables_ptr->sbr_rand_ph;
  unpack_env_prm(side + XAAC_SIDE_ENV, &c->h, &c->fb, &c->fd.f, &c->pv);
  const int16_t *hf = side + XAAC_SIDE_HF, *misc = st + XAAC_SBR_ST_MISC;
  unpack_hf_settings(hf, &c->set);
  c->fb.num_if_bands = hf[XAAC_HF_NUM_IF_BANDS];
  for (int i = 0; i < MAX_NUM_NOISE_VALUES; i++) {
    c->fd.f.sbr_invf_mode[i] = hf[XAAC_HF_INVF + i];
    c->pv.sbr_invf_mode[i] = misc[XAAC_SBR_MISC_INVF_PREV + i];
  }
  c->pv.max_qmf_subband_aac = misc[XAAC_SBR_MISC_MAX_QMF_PREV];
  c->pv.end_position = misc[XAAC_SBR_MISC_END_POS_PREV];
  ia_sbr_dec_struct *d = &c->d;
  ia_sbr_qmf_filter_bank_struct *a = &d->str_codec_qmf_bank, *s = &d->str_synthesis_qmf_bank;
  memcpy(c->anal_states, st + XAAC_SBR_ST_ANAL_STATES, sizeof(c->anal_states));
  a->no_channels = 32; a->num_time_slots = 32; a->lsb = 0; a->usb = misc[XAAC_SBR_MISC_CODEC_USB];
  a->anal_filter_states = c->anal_states;
  a->core_samples_buffer = c->anal_states + st[XAAC_SBR_ST_ANAL_POS];
  a->analy_win_coeff = qt->qmf_c;
  a->filter_pos = (WORD16 *)qt->qmf_c + st[XAAC_SBR_ST_ANAL_POS + 1];
  memcpy(c->syn_states, st + XAAC_SBR_ST_SYN_STATES, sizeof(c->syn_states));
  s->no_channels = 64; s->num_time_slots = 32; s->lsb = misc[XAAC_SBR_MISC_SYN_LSB]; s->usb = misc[XAAC_SBR_MISC_SYN_USB];
  s->filter_states = c->syn_states;
  s->ixheaacd_drc_offset = st[XAAC_SBR_ST_SYN_POS];
  s->p_filter = qt->qmf_c;
  s->filter_pos_syn = (WORD16 *)qt->qmf_c + st[XAAC_SBR_ST_SYN_POS + 1];
  unpack_sf(st + XAAC_SBR_ST_SF, &d->str_sbr_scale_fact);
  d->str_sbr_calc_env.filt_buf_me = c->filt_me;
  d->str_sbr_calc_env.filt_buf_noise_m = c->filt_noise;
  unpack_env_state(st + XAAC_SBR_ST_ENV, &d->str_sbr_calc_env);
  d->str_hf_generator.pstr_settings = &c->set;
  memcpy(d->str_hf_generator.bw_array_prev, st + XAAC_SBR_ST_BW_PREV, 6 * sizeof(int32_t));
  const int32_t *lpc = (const int32_t *)(st + XAAC_SBR_ST_LPC);
  for (int i = 0; i < 2; i++) {
    memcpy(c->lpc_r[i], lpc + 128 * i, 256); memcpy(c->lpc_i[i], lpc + 128 * i + 64, 256);
    d->str_hf_generator.lpc_filt_states_real[i] = c->lpc_r[i];
    d->str_hf_generator.lpc_filt_states_imag[i] = c->lpc_i[i];
    if (low_pow) {
      memcpy(c->lpc_lp[i], lpc + 128 * i, 128);
      d->str_hf_generator.lpc_filt_states_real[i] = c->lpc_lp[i];
    }
  }
  memcpy(c->ov, st + XAAC_SBR_ST_OV, sizeof(c->ov));
  d->ptr_sbr_overlap_buf = c->ov;
  for (int i = 0; i < 32; i++) { d->p_arr_qmf_buf_real[i] = c->qmf_out[i]; d->p_arr_qmf_buf_imag[i] = c->qmf_out[i] + 64; }
  d->pvc_qmf_enrg_arr[0] = 0;
  if (p) {
    ia_ps_dec_struct *ps = &c->ps;
    unpack_side_ps(side, ps);
    memcpy(c->ps_ap, p + XAAC_PS_ST_AP, sizeof(c->ps_ap)); memcpy(c->ps_ld, p + XAAC_PS_ST_LD, sizeof(c->ps_ld));
    memcpy(c->ps_sd, p + XAAC_PS_ST_SD, sizeof(c->ps_sd)); memcpy(c->ps_ser, p + XAAC_PS_ST_SER, sizeof(c->ps_ser));
    ps->delay_buf_qmf_ap_re_im = c->ps_ap; ps->delay_buf_qmf_ld_re_im = c->ps_ld;
    ps->delay_buf_qmf_sd_re_im = (VOID *)c->ps_sd; ps->delay_buf_qmf_ser_re_im = (VOID *)c->ps_ser;
    memcpy(ps->delay_buf_qmf_sub_re_im, p + XAAC_PS_ST_SUB, 128);
    memcpy(ps->delay_buf_qmf_sub_ser_re_im, p + XAAC_PS_ST_SUB_SER, 960);
    const int16_t *hv = p + XAAC_PS_ST_HVEC;
    memcpy(ps->h11_h12_vec, hv, 96); memcpy(ps->h21_h22_vec, hv + 48, 96); memcpy(ps->H11_H12, hv + 96, 96);
    memcpy(ps->H21_H22, hv + 144, 96); memcpy(ps->delta_h11_h12, hv + 192, 96); memcpy(ps->delta_h21_h22, hv + 240, 96);
    const int16_t *idx = p + XAAC_PS_ST_IDX;
    for (int i = 0; i < 3; i++) {
      ps->delay_buf_idx_ser[i] = idx[XAAC_PS_IDX_SER + i];
      ps->delay_sample_ser[i] = c->tabs.ps_tables_ptr->rev_link_delay_ser[i];
    }
    ps->delay_buf_idx = idx[XAAC_PS_IDX_DELAY]; ps->delay_buf_idx_long = idx[XAAC_PS_IDX_DELAY_LONG];
    ps->delay_buffer_scale = idx[XAAC_PS_IDX_SCALE]; ps->usb = idx[XAAC_PS_IDX_USB];
    memcpy(c->ps_peak, p + XAAC_PS_ST_PEAK, sizeof(c->ps_peak));
    ps->peak_decay_diff = c->ps_peak; ps->energy_prev = c->ps_peak + 20; ps->peak_decay_diff_prev = c->ps_peak + 40;
    ps->ptr_hyb_left_re = c->ps_hyb_io; ps->ptr_hyb_left_im = c->ps_hyb_io + 16;
    ps->ptr_hyb_right_re = c->ps_hyb_io + 32; ps->ptr_hyb_right_im = c->ps_hyb_io + 48;
    memcpy(c->ps_qbuf, p + XAAC_PS_ST_HYB, sizeof(c->ps_qbuf));
    ps->str_hybrid.ptr_resol = c->tabs.ps_tables_ptr->hyb_resol;
    ps->str_hybrid.ptr_qmf_buf = 12;
    ps->str_hybrid.ptr_temp_re = c->ps_temp; ps->str_hybrid.ptr_temp_im = c->ps_temp + 8;
    ps->str_hybrid.ptr_work_re = c->ps_work; ps->str_hybrid.ptr_work_im = c->ps_work + 16;
    for (int b = 0; b < 3; b++) { ps->str_hybrid.ptr_qmf_buf_re[b] = c->ps_qbuf[b][0]; ps->str_hybrid.ptr_qmf_buf_im[b] = c->ps_qbuf[b][1]; }
    ia_sbr_qmf_filter_bank_struct *r = &c->bank_r;
    memcpy(c->syn_states_r, p + XAAC_PS_ST_SYN_STATES_R, sizeof(c->syn_states_r));
    r->no_channels = 64; r->num_time_slots = 32; r->lsb = idx[XAAC_PS_IDX_LSB_R]; r->usb = idx[XAAC_PS_IDX_USB_R];
    r->filter_states = c->syn_states_r;
    r->ixheaacd_drc_offset = p[XAAC_PS_ST_SYN_POS_R];
    r->p_filter = qt->qmf_c;
    r->filter_pos_syn = (WORD16 *)qt->qmf_c + p[XAAC_PS_ST_SYN_POS_R + 1];
    unpack_sf(p + XAAC_PS_ST_SF_R, &c->sf_r);
  }
}
#endif
