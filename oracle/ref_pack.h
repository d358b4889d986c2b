/*
 * oracle/ref_pack.h — TEST INFRASTRUCTURE ONLY.
 * Conversions between the reference's pointer-rich SBR structs and the flat WORD16 records of the C-ABI
 * (include/xaac_b200.h, XAAC_ENV_* / XO_ENV_* offsets).  Used by the taps (struct -> record) and by the shim that
 * drives the compiled reference from a record (record -> struct).  OUR code, compiled against the reference headers.
 */
#ifndef XAAC_REF_PACK_H
#define XAAC_REF_PACK_H
#include "ref_headers.h"
#include "ixheaacd_env_calc.h"
#include "ixheaac_sbr_const.h"
#include "ixheaacd_pvc_dec.h"
#include "ixheaacd_sbr_dec.h"
#include "src/xaac_oracle.h"

static void pack_env_prm(int16_t *prm, const ia_sbr_header_data_struct *h, const ia_sbr_frame_info_data_struct *f,
                         const ia_sbr_prev_frame_data_struct *pv) {
  const ia_freq_band_data_struct *fb = h->pstr_freq_band_data;
  const ia_frame_info_struct *fi = &f->str_frame_info_details;
  memset(prm, 0, XO_ENV_PRM_WORDS * sizeof(int16_t));
  prm[XO_ENV_NUM_TIME_SLOTS] = h->num_time_slots;
  prm[XO_ENV_TIME_STEP] = h->time_step;
  prm[XO_ENV_CHANNEL_MODE] = (int16_t)h->channel_mode;
  prm[XO_ENV_LIMITER_GAINS] = h->limiter_gains;
  prm[XO_ENV_INTERPOL_FREQ] = h->interpol_freq;
  prm[XO_ENV_SMOOTHING_MODE] = h->smoothing_mode;
  prm[XO_ENV_NUM_SF_LO] = fb->num_sf_bands[0];
  prm[XO_ENV_NUM_SF_HI] = fb->num_sf_bands[1];
  prm[XO_ENV_NUM_NF_BANDS] = fb->num_nf_bands;
  prm[XO_ENV_SUB_BAND_START] = fb->sub_band_start;
  prm[XO_ENV_SUB_BAND_END] = fb->sub_band_end;
  prm[XO_ENV_NUM_LF_BANDS] = fb->num_lf_bands;
  prm[XO_ENV_NUM_ENV] = fi->num_env;
  prm[XO_ENV_TRANSIENT_ENV] = fi->transient_env;
  prm[XO_ENV_MAX_QMF_SUBBAND] = (int16_t)f->max_qmf_subband_aac;
  prm[XO_ENV_MAX_QMF_SUBBAND_PREV] = (int16_t)pv->max_qmf_subband_aac;
  for (int i = 0; i < MAX_ENVELOPES + 1; i++) prm[XO_ENV_BORDER_VEC + i] = fi->border_vec[i];
  for (int i = 0; i < MAX_ENVELOPES; i++) prm[XO_ENV_FREQ_RES + i] = fi->freq_res[i];
  for (int i = 0; i < MAX_NOISE_ENVELOPES + 1; i++) prm[XO_ENV_NOISE_BORDER_VEC + i] = fi->noise_border_vec[i];
  for (int i = 0; i < MAX_NUM_LIMITERS + 1; i++) prm[XO_ENV_LIM_TBL + i] = fb->freq_band_tbl_lim[i];
  for (int i = 0; i < MAX_FREQ_COEFFS / 2 + 1; i++) prm[XO_ENV_FREQ_LO + i] = fb->freq_band_tbl_lo[i];
  for (int i = 0; i < MAX_FREQ_COEFFS + 1; i++) prm[XO_ENV_FREQ_HI + i] = fb->freq_band_tbl_hi[i];
  for (int i = 0; i < MAX_NOISE_COEFFS + 1; i++) prm[XO_ENV_FREQ_NOISE + i] = fb->freq_band_tbl_noise[i];
  for (int i = 0; i < MAX_NUM_NOISE_VALUES; i++) prm[XO_ENV_NOISE_FLOOR + i] = f->int_noise_floor[i];
  for (int i = 0; i < MAX_FREQ_COEFFS; i++) prm[XO_ENV_ADD_HARMONICS + i] = (int16_t)f->add_harmonics[i];
  for (int i = 0; i < MAX_NUM_ENVELOPE_VALUES; i++) prm[XO_ENV_SF_ARR + i] = f->int_env_sf_arr[i];
}

static void unpack_env_prm(const int16_t *prm, ia_sbr_header_data_struct *h, ia_freq_band_data_struct *fb,
                           ia_sbr_frame_info_data_struct *f, ia_sbr_prev_frame_data_struct *pv) {
  ia_frame_info_struct *fi = &f->str_frame_info_details;
  h->pstr_freq_band_data = fb;
  h->num_time_slots = prm[XO_ENV_NUM_TIME_SLOTS];
  h->time_step = prm[XO_ENV_TIME_STEP];
  h->channel_mode = prm[XO_ENV_CHANNEL_MODE];
  h->limiter_gains = prm[XO_ENV_LIMITER_GAINS];
  h->interpol_freq = prm[XO_ENV_INTERPOL_FREQ];
  h->smoothing_mode = prm[XO_ENV_SMOOTHING_MODE];
  fb->num_sf_bands[0] = prm[XO_ENV_NUM_SF_LO];
  fb->num_sf_bands[1] = prm[XO_ENV_NUM_SF_HI];
  fb->num_nf_bands = prm[XO_ENV_NUM_NF_BANDS];
  fb->sub_band_start = prm[XO_ENV_SUB_BAND_START];
  fb->sub_band_end = prm[XO_ENV_SUB_BAND_END];
  fb->num_lf_bands = prm[XO_ENV_NUM_LF_BANDS];
  fb->freq_band_table[0] = fb->freq_band_tbl_lo;
  fb->freq_band_table[1] = fb->freq_band_tbl_hi;
  fi->num_env = prm[XO_ENV_NUM_ENV];
  fi->transient_env = prm[XO_ENV_TRANSIENT_ENV];
  f->max_qmf_subband_aac = prm[XO_ENV_MAX_QMF_SUBBAND];
  pv->max_qmf_subband_aac = prm[XO_ENV_MAX_QMF_SUBBAND_PREV];
  for (int i = 0; i < MAX_ENVELOPES + 1; i++) fi->border_vec[i] = prm[XO_ENV_BORDER_VEC + i];
  for (int i = 0; i < MAX_ENVELOPES; i++) fi->freq_res[i] = prm[XO_ENV_FREQ_RES + i];
  for (int i = 0; i < MAX_NOISE_ENVELOPES + 1; i++) fi->noise_border_vec[i] = prm[XO_ENV_NOISE_BORDER_VEC + i];
  for (int i = 0; i < MAX_NUM_LIMITERS + 1; i++) fb->freq_band_tbl_lim[i] = prm[XO_ENV_LIM_TBL + i];
  for (int i = 0; i < MAX_FREQ_COEFFS / 2 + 1; i++) fb->freq_band_tbl_lo[i] = prm[XO_ENV_FREQ_LO + i];
  for (int i = 0; i < MAX_FREQ_COEFFS + 1; i++) fb->freq_band_tbl_hi[i] = prm[XO_ENV_FREQ_HI + i];
  for (int i = 0; i < MAX_NOISE_COEFFS + 1; i++) fb->freq_band_tbl_noise[i] = prm[XO_ENV_FREQ_NOISE + i];
  for (int i = 0; i < MAX_NUM_NOISE_VALUES; i++) f->int_noise_floor[i] = prm[XO_ENV_NOISE_FLOOR + i];
  for (int i = 0; i < MAX_FREQ_COEFFS; i++) f->add_harmonics[i] = prm[XO_ENV_ADD_HARMONICS + i];
  for (int i = 0; i < MAX_NUM_ENVELOPE_VALUES; i++) f->int_env_sf_arr[i] = prm[XO_ENV_SF_ARR + i];
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
  memset(st, 0, XO_ENV_ST_WORDS * sizeof(int16_t));
  memcpy(st + XO_ENV_ST_FILT_ME, e->filt_buf_me, 2 * MAX_FREQ_COEFFS * sizeof(int16_t));
  memcpy(st + XO_ENV_ST_FILT_NOISE, e->filt_buf_noise_m, MAX_FREQ_COEFFS * sizeof(int16_t));
  st[XO_ENV_ST_NOISE_E] = (int16_t)e->filt_buf_noise_e;
  st[XO_ENV_ST_START_UP] = (int16_t)e->start_up;
  st[XO_ENV_ST_PH_INDEX] = e->ph_index;
  st[XO_ENV_ST_TRANS_PREV] = e->tansient_env_prev;
  st[XO_ENV_ST_HARM_INDEX] = e->harm_index;
  for (int i = 0; i < MAX_FREQ_COEFFS; i++) st[XO_ENV_ST_HARM_PREV + i] = e->harm_flags_prev[i];
}
/* e->filt_buf_me / filt_buf_noise_m must already point at caller-owned buffers */
static void unpack_env_state(const int16_t *st, ia_sbr_calc_env_struct *e) {
  memcpy(e->filt_buf_me, st + XO_ENV_ST_FILT_ME, 2 * MAX_FREQ_COEFFS * sizeof(int16_t));
  memcpy(e->filt_buf_noise_m, st + XO_ENV_ST_FILT_NOISE, MAX_FREQ_COEFFS * sizeof(int16_t));
  e->filt_buf_noise_e = st[XO_ENV_ST_NOISE_E];
  e->start_up = st[XO_ENV_ST_START_UP];
  e->ph_index = st[XO_ENV_ST_PH_INDEX];
  e->tansient_env_prev = st[XO_ENV_ST_TRANS_PREV];
  e->harm_index = st[XO_ENV_ST_HARM_INDEX];
  for (int i = 0; i < MAX_FREQ_COEFFS; i++) e->harm_flags_prev[i] = (WORD8)st[XO_ENV_ST_HARM_PREV + i];
}
#endif
