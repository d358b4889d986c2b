/*
 * oracle/ref_taps_esbr.c — TEST INFRASTRUCTURE ONLY.
 *
 * Stage taps for the float eSBR HF generator and envelope adjuster of the UNMODIFIED reference decoder, installed with
 * `ld --wrap=ixheaacd_generate_hf --wrap=ixheaacd_sbr_env_calc` into oracle/_ref/xaacdec_tap (mechanism: oracle/ref_taps.c).
 * Records are written in the flat XO_EHF_* / XO_EEC_* layouts of oracle/src/xaac_oracle.h:
 *  <tap>.ehf: int32 'EHF1', int32 ret, int32 has_pv, int32 ldmps, int32 par[96], float bw_in[6], bw_out[6], int32 patch_out[8], patch_in[8],
 *             float src_re[2560], src_im[2560], pv_re[2560], pv_im[2560], dst_in_re, dst_in_im, dst_out_re, dst_out_im
 *  <tap>.eec: int32 'EEC1', int32 ret, int32 ldmps, int32 ipar_in[288], ipar_out[288], float fpar[464], state_in[640],
 *             state_out[640], re_in[2560], im_in[2560], re_out[2560], im_out[2560]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#define REF_SHIM_HEADERS_ONLY
#include "ref_headers.h"
#include "src/xaac_oracle.h"

static FILE *open_tap(const char *stage, const char *ext) {
  const char *p = getenv("XAAC_TAP_FILE"), *s = getenv("XAAC_TAP_STAGES");
  if (p && *p && s && strstr(s, stage)) {
    char name[1024];
    snprintf(name, sizeof(name), "%s.%s", p, ext);
    return fopen(name, "wb");
  }
  return NULL;
}
static int tap_max(void) {
  const char *m = getenv("XAAC_TAP_MAX");
  return m ? atoi(m) : 1000000;
}

WORD32 __real_ixheaacd_generate_hf(FLOAT32 a[][64], FLOAT32 b[][64], FLOAT32 c[][64], FLOAT32 d[][64], FLOAT32 e[][64],
                                   FLOAT32 f[][64], ia_sbr_frame_info_data_struct *fd, ia_sbr_header_data_struct *hd,
                                   WORD32 ldmps, WORD32 time_slots, WORD32 ec_flag);
WORD32 __wrap_ixheaacd_generate_hf(FLOAT32 src_re[][64], FLOAT32 src_im[][64], FLOAT32 pv_re[][64], FLOAT32 pv_im[][64],
                                   FLOAT32 dst_re[][64], FLOAT32 dst_im[][64], ia_sbr_frame_info_data_struct *fd,
                                   ia_sbr_header_data_struct *hd, WORD32 ldmps, WORD32 time_slots, WORD32 ec_flag) {
  static FILE *fp = NULL;
  static int tried = 0, count = 0;
  if (!tried) {
    tried = 1;
    fp = open_tap("ehf", "ehf");
  }
  const int rec = fp && count < tap_max();
  static float din[2][2560];
  int32_t head[4], par[XO_EHF_PAR_WORDS], patch[8], patch_in[8];
  float bw_in[6];
  const int has_pv = hd->hbe_flag && pv_re && pv_im;
  if (rec) {
    ia_freq_band_data_struct *fb = hd->pstr_freq_band_data;
    memset(par, 0, sizeof(par));
    par[XO_EHF_NUM_MF] = fb->num_mf_bands;
    par[XO_EHF_NUM_IF] = fb->num_nf_bands;
    par[XO_EHF_SB_START] = fb->sub_band_start;
    par[XO_EHF_BORDER_FIRST] = fd->str_frame_info_details.border_vec[0];
    par[XO_EHF_BORDER_LAST] = fd->str_frame_info_details.border_vec[fd->str_frame_info_details.num_env];
    par[XO_EHF_HBE_FLAG] = hd->hbe_flag;
    par[XO_EHF_PATCHING_MODE] = fd->sbr_patching_mode;
    par[XO_EHF_FS] = hd->out_sampling_freq;
    par[XO_EHF_PRE_PROC] = hd->pre_proc_flag;
    par[XO_EHF_USF4] = hd->is_usf_4;
    par[XO_EHF_MPS_SBR] = fd->mps_sbr_flag;
    par[XO_EHF_COV_COUNT] = fd->cov_count;
    for (int i = 0; i < 5; i++) {
      par[XO_EHF_INVF + i] = fd->sbr_invf_mode[i];
      par[XO_EHF_INVF_PREV + i] = fd->sbr_invf_mode_prev[i];
      par[XO_EHF_INVF_TBL + i] = fb->freq_band_tbl_noise[1 + i];
    }
    for (int i = 0; i < 57; i++) par[XO_EHF_FMASTER + i] = fb->f_master_tbl[i];
    for (int i = 0; i < 6; i++) bw_in[i] = fd->bw_array_prev[i];
    patch_in[0] = fd->patch_param.num_patches;
    for (int i = 0; i < 7; i++) patch_in[1 + i] = fd->patch_param.start_subband[i];
    memcpy(din[0], dst_re - 2, sizeof(din[0]));
    memcpy(din[1], dst_im - 2, sizeof(din[1]));
  }
  WORD32 ret = __real_ixheaacd_generate_hf(src_re, src_im, pv_re, pv_im, dst_re, dst_im, fd, hd, ldmps, time_slots, ec_flag);
  if (rec) {
    static float zero[2560];
    head[0] = 0x31464845;
    head[1] = ret;
    head[2] = has_pv;
    head[3] = ldmps;
    patch[0] = fd->patch_param.num_patches;
    for (int i = 0; i < 7; i++) patch[1 + i] = fd->patch_param.start_subband[i];
    fwrite(head, 4, 4, fp);
    fwrite(par, 4, XO_EHF_PAR_WORDS, fp);
    fwrite(bw_in, 4, 6, fp);
    fwrite(fd->bw_array_prev, 4, 6, fp);
    fwrite(patch, 4, 8, fp);
    fwrite(patch_in, 4, 8, fp);
    fwrite(src_re - 2, 4, 2560, fp);
    fwrite(src_im - 2, 4, 2560, fp);
    fwrite(has_pv ? (float *)(pv_re - 2) : zero, 4, 2560, fp);
    fwrite(has_pv ? (float *)(pv_im - 2) : zero, 4, 2560, fp);
    fwrite(din[0], 4, 2560, fp);
    fwrite(din[1], 4, 2560, fp);
    fwrite(dst_re - 2, 4, 2560, fp);
    fwrite(dst_im - 2, 4, 2560, fp);
    fflush(fp);
    count++;
  }
  return ret;
}

WORD32 __real_ixheaacd_sbr_env_calc(ia_sbr_frame_info_data_struct *fd, FLOAT32 a[][64], FLOAT32 b[][64], FLOAT32 c[][64],
                                    FLOAT32 d[][64], WORD32 x_over_qmf[MAX_NUM_PATCHES], FLOAT32 *scratch, FLOAT32 *env_out,
                                    WORD32 ldmps, WORD32 ec_flag);
static void eec_pack(int32_t *ip, const ia_sbr_frame_info_data_struct *fd) {
  const ia_sbr_header_data_struct *hd = fd->pstr_sbr_header;
  const ia_freq_band_data_struct *fb = hd->pstr_freq_band_data;
  const ia_frame_info_struct *fi = &fd->str_frame_info_details;
  memset(ip, 0, 4 * XO_EEC_IPAR_WORDS);
  ip[XO_EEC_SB_START] = fb->sub_band_start;
  ip[XO_EEC_SB_END] = fb->sub_band_end;
  ip[XO_EEC_NUM_ENV] = fi->num_env;
  ip[XO_EEC_TRANS_ENV] = fi->transient_env;
  ip[XO_EEC_SHORT_PREV] = fd->env_short_flag_prev;
  ip[XO_EEC_NUM_NOISE_ENV] = fi->num_noise_env;
  ip[XO_EEC_NUM_SF_LO] = fb->num_sf_bands[0];
  ip[XO_EEC_NUM_SF_HI] = fb->num_sf_bands[1];
  ip[XO_EEC_NUM_NF] = fb->num_nf_bands;
  ip[XO_EEC_SMOOTHING_MODE] = hd->smoothing_mode;
  ip[XO_EEC_INTERPOL_FREQ] = hd->interpol_freq;
  ip[XO_EEC_LIMITER_BANDS] = hd->limiter_bands;
  ip[XO_EEC_LIMITER_GAINS] = hd->limiter_gains;
  ip[XO_EEC_HARM_INDEX] = fd->harm_index;
  ip[XO_EEC_PHASE_INDEX] = fd->phase_index;
  ip[XO_EEC_START_UP] = hd->esbr_start_up;
  ip[XO_EEC_RESET] = fd->reset_flag;
  ip[XO_EEC_SBR_MODE] = fd->sbr_mode;
  ip[XO_EEC_USF4] = hd->is_usf_4;
  ip[XO_EEC_PATCHING_CHANGED] = fd->sbr_patching_mode != fd->prev_sbr_patching_mode;
  for (int i = 0; i < 9; i++) ip[XO_EEC_BORDER + i] = fi->border_vec[i];
  for (int i = 0; i < 8; i++) ip[XO_EEC_FREQ_RES + i] = fi->freq_res[i];
  for (int i = 0; i < 3; i++) ip[XO_EEC_NOISE_BORDER + i] = fi->noise_border_vec[i];
  for (int i = 0; i < 8; i++) ip[XO_EEC_INTER_TES + i] = fd->inter_temp_shape_mode[i];
  for (int i = 0; i < 4; i++) ip[XO_EEC_GATE_MODE + i] = fd->gate_mode[i];
  for (int i = 0; i < 52; i++) ip[XO_EEC_LIM_TABLE + i] = fd->lim_table[i / 13][i % 13];
  for (int i = 0; i < 6; i++) ip[XO_EEC_TBL_NOISE + i] = fb->freq_band_tbl_noise[i];
  for (int i = 0; i < 29; i++) ip[XO_EEC_TBL_LO + i] = fb->freq_band_tbl_lo[i];
  for (int i = 0; i < 57; i++) ip[XO_EEC_TBL_HI + i] = fb->freq_band_tbl_hi[i];
  for (int i = 0; i < 56; i++) ip[XO_EEC_ADD_HARM + i] = fd->add_harmonics[i];
  memcpy(ip + XO_EEC_HARM_PREV, fd->harm_flag_prev, 64);
}
WORD32 __wrap_ixheaacd_sbr_env_calc(ia_sbr_frame_info_data_struct *fd, FLOAT32 re[][64], FLOAT32 im[][64], FLOAT32 re1[][64],
                                    FLOAT32 im1[][64], WORD32 x_over_qmf[MAX_NUM_PATCHES], FLOAT32 *scratch,
                                    FLOAT32 *env_out, WORD32 ldmps, WORD32 ec_flag) {
  static FILE *fp = NULL;
  static int tried = 0, count = 0;
  if (!tried) {
    tried = 1;
    fp = open_tap("eec", "eec");
  }
  const int rec = fp && count < tap_max();
  static int32_t ip_in[XO_EEC_IPAR_WORDS], ip_out[XO_EEC_IPAR_WORDS];
  static float fpar[XO_EEC_FPAR_WORDS], st_in[640], qin[2][2560];
  if (rec) {
    eec_pack(ip_in, fd);
    memset(fpar, 0, sizeof(fpar));
    memcpy(fpar + XO_EEC_SFB_NRG, fd->flt_env_sf_arr, 448 * 4);
    memcpy(fpar + XO_EEC_NOISE_FLOOR, fd->flt_noise_floor, 10 * 4);
    memcpy(st_in, fd->e_gain, 320 * 4);
    memcpy(st_in + 320, fd->noise_buf, 320 * 4);
    memcpy(qin[0], re - 2, sizeof(qin[0]));
    memcpy(qin[1], im - 2, sizeof(qin[1]));
  }
  WORD32 ret = __real_ixheaacd_sbr_env_calc(fd, re, im, re1, im1, x_over_qmf, scratch, env_out, ldmps, ec_flag);
  if (rec) {
    int32_t head[3] = {0x31434545, ret, ldmps};
    eec_pack(ip_out, fd);
    fwrite(head, 4, 3, fp);
    fwrite(ip_in, 4, XO_EEC_IPAR_WORDS, fp);
    fwrite(ip_out, 4, XO_EEC_IPAR_WORDS, fp);
    fwrite(fpar, 4, XO_EEC_FPAR_WORDS, fp);
    fwrite(st_in, 4, 640, fp);
    fwrite(fd->e_gain, 4, 320, fp);
    fwrite(fd->noise_buf, 4, 320, fp);
    fwrite(qin[0], 4, 2560, fp);
    fwrite(qin[1], 4, 2560, fp);
    fwrite(re - 2, 4, 2560, fp);
    fwrite(im - 2, 4, 2560, fp);
    fflush(fp);
    count++;
  }
  return ret;
}
