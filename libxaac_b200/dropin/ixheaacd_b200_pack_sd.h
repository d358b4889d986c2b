/*
 * libxaac_b200/dropin/ixheaacd_b200_pack_sd.h — reference-side half of the SBR side-info dequantisation hand-over
 * (ixheaacd_dec_sbrdata, decoder/ixheaacd_env_dec.c:628): header / frame-data / previous-frame structs <-> XAAC_SD_* record
 * (include/xaac_b200.h).  Every record word is the reference member of the same name.  Used by ixheaacd_b200_glue.c and by
 * oracle/ref_shim_sd.c.
 */
#ifndef IXHEAACD_B200_PACK_SD_H
#define IXHEAACD_B200_PACK_SD_H
#include <string.h>
#include "ixheaacd_b200_ref_headers.h"
#include "xaac_b200.h"

static void b200_sd_pack_ch(int16_t *b, const ia_sbr_header_data_struct *h, const ia_sbr_frame_info_data_struct *f,
                            const ia_sbr_prev_frame_data_struct *p) {
  const ia_frame_info_struct *fi = &f->str_frame_info_details;
  int i;
  b[XAAC_SDC_NUM_SF_LO] = h->pstr_freq_band_data->num_sf_bands[0];
  b[XAAC_SDC_NUM_SF_HI] = h->pstr_freq_band_data->num_sf_bands[1];
  b[XAAC_SDC_NUM_NF] = h->pstr_freq_band_data->num_nf_bands;
  b[XAAC_SDC_NUM_TIME_SLOTS] = h->num_time_slots;
  b[XAAC_SDC_ERR_FLAG] = (int16_t)h->err_flag;
  b[XAAC_SDC_ERR_FLAG_PREV] = (int16_t)h->err_flag_prev;
  b[XAAC_SDC_HDR_AMP_RES] = h->amp_res;
  b[XAAC_SDC_NUM_NOISE_SFAC] = (int16_t)f->num_noise_sfac;
  b[XAAC_SDC_NUM_ENV] = fi->num_env;
  b[XAAC_SDC_NUM_NOISE_ENV] = fi->num_noise_env;
  b[XAAC_SDC_TRANSIENT_ENV] = fi->transient_env;
  b[XAAC_SDC_AMP_RES] = f->amp_res;
  b[XAAC_SDC_COUPLING] = (int16_t)f->coupling_mode;
  b[XAAC_SDC_NUM_ENV_SFAC] = f->num_env_sfac;
  b[XAAC_SDC_MAX_QMF_SB] = (int16_t)f->max_qmf_subband_aac;
  memcpy(b + XAAC_SDC_FREQ_RES, fi->freq_res, 8 * 2);
  memcpy(b + XAAC_SDC_BORDER, fi->border_vec, 9 * 2);
  memcpy(b + XAAC_SDC_NOISE_BORDER, fi->noise_border_vec, 3 * 2);
  memcpy(b + XAAC_SDC_DIR, f->del_cod_dir_arr, 8 * 2);
  memcpy(b + XAAC_SDC_DIR_NOISE, f->del_cod_dir_noise_arr, 2 * 2);
  for (i = 0; i < 10; i++) b[XAAC_SDC_INVF + i] = (int16_t)f->sbr_invf_mode[i];
  for (i = 0; i < 56; i++) b[XAAC_SDC_ADD_HARM + i] = (int16_t)f->add_harmonics[i];
  memcpy(b + XAAC_SDC_ENV, f->int_env_sf_arr, 448 * 2);
  memcpy(b + XAAC_SDC_NOISE, f->int_noise_floor, 10 * 2);
  memcpy(b + XAAC_SDC_PREV_NRG, p->sfb_nrg_prev, 56 * 2);
  memcpy(b + XAAC_SDC_PREV_NOISE, p->prev_noise_level, 5 * 2);
  b[XAAC_SDC_PREV_AMP_RES] = p->amp_res;
  b[XAAC_SDC_PREV_END_POS] = p->end_position;
  b[XAAC_SDC_PREV_MAX_QMF] = (int16_t)p->max_qmf_subband_aac;
  b[XAAC_SDC_PREV_COUPLING] = (int16_t)p->coupling_mode;
  for (i = 0; i < 10; i++) b[XAAC_SDC_PREV_INVF + i] = (int16_t)p->sbr_invf_mode[i];
}

static void b200_sd_unpack_ch(const int16_t *b, ia_sbr_header_data_struct *h, ia_sbr_frame_info_data_struct *f,
                              ia_sbr_prev_frame_data_struct *p) {
  ia_frame_info_struct *fi = &f->str_frame_info_details;
  int i;
  h->err_flag = b[XAAC_SDC_ERR_FLAG];
  h->err_flag_prev = b[XAAC_SDC_ERR_FLAG_PREV];
  f->num_noise_sfac = b[XAAC_SDC_NUM_NOISE_SFAC];
  fi->num_env = b[XAAC_SDC_NUM_ENV];
  fi->num_noise_env = b[XAAC_SDC_NUM_NOISE_ENV];
  fi->transient_env = b[XAAC_SDC_TRANSIENT_ENV];
  f->amp_res = b[XAAC_SDC_AMP_RES];
  f->coupling_mode = b[XAAC_SDC_COUPLING];
  f->num_env_sfac = b[XAAC_SDC_NUM_ENV_SFAC];
  f->max_qmf_subband_aac = b[XAAC_SDC_MAX_QMF_SB];
  memcpy(fi->freq_res, b + XAAC_SDC_FREQ_RES, 8 * 2);
  memcpy(fi->border_vec, b + XAAC_SDC_BORDER, 9 * 2);
  memcpy(fi->noise_border_vec, b + XAAC_SDC_NOISE_BORDER, 3 * 2);
  memcpy(f->del_cod_dir_arr, b + XAAC_SDC_DIR, 8 * 2);
  memcpy(f->del_cod_dir_noise_arr, b + XAAC_SDC_DIR_NOISE, 2 * 2);
  for (i = 0; i < 10; i++) f->sbr_invf_mode[i] = b[XAAC_SDC_INVF + i];
  for (i = 0; i < 56; i++) f->add_harmonics[i] = b[XAAC_SDC_ADD_HARM + i];
  memcpy(f->int_env_sf_arr, b + XAAC_SDC_ENV, 448 * 2);
  memcpy(f->int_noise_floor, b + XAAC_SDC_NOISE, 10 * 2);
  memcpy(p->sfb_nrg_prev, b + XAAC_SDC_PREV_NRG, 56 * 2);
  memcpy(p->prev_noise_level, b + XAAC_SDC_PREV_NOISE, 5 * 2);
}

/* 0, or -1 when the call is outside what the device stage covers (the caller then runs the reference's own code) */
static int b200_sd_pack(int16_t *rec, ia_sbr_header_data_struct *h0, ia_sbr_header_data_struct *h1,
                        ia_sbr_frame_info_data_struct *f0, ia_sbr_prev_frame_data_struct *p0, ia_sbr_frame_info_data_struct *f1,
                        ia_sbr_prev_frame_data_struct *p1, WORD32 ldmps_present, WORD32 audio_object_type, WORD32 ec_flag) {
  if (ldmps_present || ec_flag || audio_object_type == AOT_ER_AAC_ELD || h0->usac_flag || h0->enh_sbr) return -1;
  if (f1 && (!h1 || !p1 || h1->usac_flag || h1->enh_sbr)) return -1;
  memset(rec, 0, XAAC_SD_WORDS * 2);
  rec[XAAC_SD_NUM_CH] = f1 ? 2 : 1;
  rec[XAAC_SD_SHARED_HDR] = (f1 && h0 == h1) ? 1 : 0;
  b200_sd_pack_ch(rec + XAAC_SD_CH, h0, f0, p0);
  if (f1) b200_sd_pack_ch(rec + XAAC_SD_CH + XAAC_SD_CH_WORDS, h1, f1, p1);
  return 0;
}
static void b200_sd_unpack(const int16_t *rec, ia_sbr_header_data_struct *h0, ia_sbr_header_data_struct *h1,
                           ia_sbr_frame_info_data_struct *f0, ia_sbr_prev_frame_data_struct *p0,
                           ia_sbr_frame_info_data_struct *f1, ia_sbr_prev_frame_data_struct *p1) {
  if (f1 && h0 != h1) b200_sd_unpack_ch(rec + XAAC_SD_CH + XAAC_SD_CH_WORDS, h1, f1, p1);
  else if (f1) {
    ia_sbr_header_data_struct tmp = *h1; /* shared header: its flags live in channel block 0 */
    b200_sd_unpack_ch(rec + XAAC_SD_CH + XAAC_SD_CH_WORDS, &tmp, f1, p1);
  }
  b200_sd_unpack_ch(rec + XAAC_SD_CH, h0, f0, p0);
}
/* ---- ixheaacd_decode_ps_data (decoder/ixheaacd_ps_bitdec.c:98): ia_ps_dec_struct <-> XAAC_PSD_* record ---- */
static void b200_psd_pack(int16_t *r, const ia_ps_dec_struct *ps, WORD32 frame_size) {
  int i;
  memset(r, 0, XAAC_PSD_WORDS * 2);
  r[XAAC_PSD_DATA_PRESENT] = (int16_t)ps->ps_data_present;
  r[XAAC_PSD_ENABLE_IID] = (int16_t)ps->enable_iid;
  r[XAAC_PSD_ENABLE_ICC] = (int16_t)ps->enable_icc;
  r[XAAC_PSD_IID_MODE] = ps->iid_mode;
  r[XAAC_PSD_ICC_MODE] = ps->icc_mode;
  r[XAAC_PSD_IID_QUANT] = (int16_t)ps->iid_quant;
  r[XAAC_PSD_FRAME_CLASS] = (int16_t)ps->frame_class;
  r[XAAC_PSD_NUM_ENV] = ps->num_env;
  r[XAAC_PSD_FRAME_SIZE] = (int16_t)frame_size;
  memcpy(r + XAAC_PSD_BORDER, ps->border_position, 7 * 2);
  for (i = 0; i < 5; i++) {
    r[XAAC_PSD_IID_DT + i] = (int16_t)ps->iid_dt[i];
    r[XAAC_PSD_ICC_DT + i] = (int16_t)ps->icc_dt[i];
  }
  memcpy(r + XAAC_PSD_IID_TABLE, ps->iid_par_table, 7 * 34 * 2);
  memcpy(r + XAAC_PSD_ICC_TABLE, ps->icc_par_table, 7 * 34 * 2);
  memcpy(r + XAAC_PSD_IID_PREV, ps->iid_par_prev, 34 * 2);
  memcpy(r + XAAC_PSD_ICC_PREV, ps->icc_par_prev, 34 * 2);
}
static void b200_psd_unpack(const int16_t *r, ia_ps_dec_struct *ps) {
  ps->ps_data_present = r[XAAC_PSD_DATA_PRESENT];
  ps->num_env = r[XAAC_PSD_NUM_ENV];
  memcpy(ps->border_position, r + XAAC_PSD_BORDER, 7 * 2);
  memcpy(ps->iid_par_table, r + XAAC_PSD_IID_TABLE, 7 * 34 * 2);
  memcpy(ps->icc_par_table, r + XAAC_PSD_ICC_TABLE, 7 * 34 * 2);
  memcpy(ps->iid_par_prev, r + XAAC_PSD_IID_PREV, 34 * 2);
  memcpy(ps->icc_par_prev, r + XAAC_PSD_ICC_PREV, 34 * 2);
}
#endif
