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
#include "ixheaacd_env_calc.h"
#include "ixheaac_sbr_const.h"
#include "ixheaacd_pvc_dec.h"
#include "ixheaacd_sbr_dec.h"
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

/* last packed records of the inner stages, for the whole-stage tap below */
static int32_t g_ehf_par[XO_EHF_PAR_WORDS], g_ehf_patch_in[8];
static float g_ehf_bw_in[6];
static int g_ehf_called, g_eec_called;
static int32_t g_eec_ipar_in[XO_EEC_IPAR_WORDS], g_eec_ipar_out[XO_EEC_IPAR_WORDS];
static float g_eec_fpar[XO_EEC_FPAR_WORDS], g_eec_state_in[640];

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
  {
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
    memcpy(g_ehf_par, par, sizeof(par));
    memcpy(g_ehf_patch_in, patch_in, sizeof(patch_in));
    memcpy(g_ehf_bw_in, bw_in, sizeof(bw_in));
    g_ehf_called++;
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
  {
    eec_pack(ip_in, fd);
    memset(fpar, 0, sizeof(fpar));
    memcpy(fpar + XO_EEC_SFB_NRG, fd->flt_env_sf_arr, 448 * 4);
    memcpy(fpar + XO_EEC_NOISE_FLOOR, fd->flt_noise_floor, 10 * 4);
    memcpy(st_in, fd->e_gain, 320 * 4);
    memcpy(st_in + 320, fd->noise_buf, 320 * 4);
    memcpy(qin[0], re - 2, sizeof(qin[0]));
    memcpy(qin[1], im - 2, sizeof(qin[1]));
    memcpy(g_eec_ipar_in, ip_in, sizeof(ip_in));
    memcpy(g_eec_fpar, fpar, sizeof(fpar));
    memcpy(g_eec_state_in, st_in, sizeof(st_in));
    g_eec_called++;
  }
  WORD32 ret = __real_ixheaacd_sbr_env_calc(fd, re, im, re1, im1, x_over_qmf, scratch, env_out, ldmps, ec_flag);
  eec_pack(ip_out, fd);
  memcpy(g_eec_ipar_out, ip_out, sizeof(ip_out));
  if (rec) {
    int32_t head[3] = {0x31434545, ret, ldmps};
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


/* ---- whole eSBR stage: hooks called from __wrap_ixheaacd_sbr_dec (oracle/ref_taps.c) around the real call -----------------
 * <tap>.esd record: int32 'ESD1', int32 head[15] = {ret, apply, hbe_flag, ps, stereo_config_idx, mps_sbr_flag, sbr_mode,
 *   qmf_sb_prev (at entry), sub_band_start, border_vec[0], sbr_ratio_idx, usac_flag, channel id, generate_hf calls, env_calc calls},
 *   float time_in[1024], state_in {float qmf_re[2560], qmf_im[2560], out_re[2560], out_im[2560], int32 anal[320], apos[2],
 *   synth[1280], spos[2], float bw[6], int32 patch[8], float ec_state[640]}, int32 hf_par[96], ec_ipar_in[288], float ec_fpar[464],
 *   float time_out[2048], state_out {same members}, int32 ec_ipar_out[288] */
typedef struct {
  float q[4][2560];
  int32_t anal[320], apos[2], synth[1280], spos[2];
  float bw[6];
  int32_t patch[8];
  float ec[640];
} esd_state_t;
static void esd_state(esd_state_t *s, ia_sbr_dec_struct *d, ia_sbr_frame_info_data_struct *f, ia_sbr_tables_struct *t) {
  ia_sbr_qmf_filter_bank_struct *a = &d->str_codec_qmf_bank, *y = &d->str_synthesis_qmf_bank;
  WORD32 *c = (WORD32 *)t->qmf_dec_tables_ptr->esbr_qmf_c;
  memcpy(s->q[0], d->qmf_buf_real, sizeof(s->q[0]));
  memcpy(s->q[1], d->qmf_buf_imag, sizeof(s->q[1]));
  memcpy(s->q[2], d->sbr_qmf_out_real, sizeof(s->q[2]));
  memcpy(s->q[3], d->sbr_qmf_out_imag, sizeof(s->q[3]));
  memcpy(s->anal, a->anal_filter_states_32, sizeof(s->anal));
  s->apos[0] = (int32_t)(a->state_new_samples_pos_low_32 - a->anal_filter_states_32);
  s->apos[1] = (int32_t)(a->filter_pos_32 - c);
  memcpy(s->synth, y->filter_states_32, sizeof(s->synth));
  s->spos[0] = y->ixheaacd_drc_offset;
  s->spos[1] = (int32_t)(y->filter_pos_syn_32 - y->p_filter_32);
  memcpy(s->bw, f->bw_array_prev, sizeof(s->bw));
  s->patch[0] = f->patch_param.num_patches;
  for (int i = 0; i < 7; i++) s->patch[1 + i] = f->patch_param.start_subband[i];
  memcpy(s->ec, f->e_gain, 320 * 4);
  memcpy(s->ec + 320, f->noise_buf, 320 * 4);
}
static FILE *esd_fp;
static int esd_count, esd_rec;
static int32_t esd_head[16];
static float esd_tin[1024];
static esd_state_t esd_in, esd_out;
static int esh_rec;
static void esh_pre(ia_sbr_dec_struct *d, ia_sbr_header_data_struct *h, ia_sbr_frame_info_data_struct *f, ia_sbr_tables_struct *t,
                    int eligible);
void esbr_stage_tap_pre(ia_sbr_dec_struct *d, ia_sbr_header_data_struct *h, ia_sbr_frame_info_data_struct *f,
                        ia_ps_dec_struct *ps, ia_sbr_tables_struct *t, int apply, int low_pow, int aot, int ldmps, int drc_on) {
  static int tried = 0;
  static void *chan[8];
  if (!tried) {
    tried = 1;
    esd_fp = open_tap("esd", "esd");
  }
  const int eligible = h->enh_sbr && h->usac_flag && !low_pow && !ldmps && !drc_on && h->num_time_slots == 16 &&
                       d->str_codec_qmf_bank.no_channels == 32 && d->str_synthesis_qmf_bank.no_channels == 64;
  esd_rec = esd_fp && esd_count < tap_max() && eligible;
  esh_pre(d, h, f, t, eligible);
  if (!esd_rec && !esh_rec) return;
  int ch = 0;
  while (ch < 7 && chan[ch] && chan[ch] != (void *)d) ch++;
  chan[ch] = d;
  g_ehf_called = g_eec_called = 0;
  esd_head[0] = 0x31445345;
  esd_head[2] = apply;
  esd_head[3] = h->hbe_flag;
  esd_head[4] = (h->channel_mode == PS_STEREO) || h->enh_sbr_ps;
  esd_head[5] = f->stereo_config_idx;
  esd_head[6] = f->mps_sbr_flag;
  esd_head[7] = f->sbr_mode;
  esd_head[8] = h->pstr_freq_band_data->qmf_sb_prev;
  esd_head[9] = h->pstr_freq_band_data->sub_band_start;
  esd_head[10] = f->str_frame_info_details.border_vec[0];
  esd_head[11] = h->sbr_ratio_idx;
  esd_head[12] = h->usac_flag;
  esd_head[13] = ch;
  memcpy(esd_tin, d->time_sample_buf, sizeof(esd_tin));
  esd_state(&esd_in, d, f, t);
}
/* ---- the same stage with the harmonic transposer (hbe_flag): <tap>.esh record
 *   int32 'ESH1', head[15] as above, float time_in[1024], esh_state_t state_in, int32 hf_par[96], ec_ipar_in[288], float ec_fpar[464],
 *   int32 hbe_cfg[16], float time_out[2048], esh_state_t state_out, int32 ec_ipar_out[288] */
typedef struct {
  float qre[72 * 64], qim[72 * 64], ore[2560], oim[2560], pre[2560], pim[2560];
  int32_t anal[320], apos[2], synth[1280], spos[2];
  float bw[6];
  int32_t patch[8];
  float ec[640];
  float hbe[XO_HBE_ST_WORDS];
} esh_state_t;
extern int32_t g_hbe_cfg[XO_HBE_CFG_WORDS];
extern int g_hbe_called;
int hbe_state_pack(float *st, const ia_esbr_hbe_txposer_struct *t);
static int esh_state(esh_state_t *s, ia_sbr_dec_struct *d, ia_sbr_frame_info_data_struct *f, ia_sbr_tables_struct *t) {
  static esd_state_t tmp;
  esd_state(&tmp, d, f, t);
  memcpy(s->qre, d->qmf_buf_real, sizeof(s->qre));
  memcpy(s->qim, d->qmf_buf_imag, sizeof(s->qim));
  memcpy(s->ore, tmp.q[2], sizeof(s->ore));
  memcpy(s->oim, tmp.q[3], sizeof(s->oim));
  memcpy(s->pre, d->ph_vocod_qmf_real, sizeof(s->pre));
  memcpy(s->pim, d->ph_vocod_qmf_imag, sizeof(s->pim));
  memcpy(s->anal, tmp.anal, sizeof(s->anal));
  memcpy(s->apos, tmp.apos, sizeof(s->apos));
  memcpy(s->synth, tmp.synth, sizeof(s->synth));
  memcpy(s->spos, tmp.spos, sizeof(s->spos));
  memcpy(s->bw, tmp.bw, sizeof(s->bw));
  memcpy(s->patch, tmp.patch, sizeof(s->patch));
  memcpy(s->ec, tmp.ec, sizeof(s->ec));
  return d->p_hbe_txposer ? hbe_state_pack(s->hbe, d->p_hbe_txposer) : -1;
}
static FILE *esh_fp;
static int esh_count, esh_bad;
static esh_state_t esh_in, esh_out;
static void esh_pre(ia_sbr_dec_struct *d, ia_sbr_header_data_struct *h, ia_sbr_frame_info_data_struct *f, ia_sbr_tables_struct *t,
                    int eligible) {
  static int tried = 0;
  if (!tried) {
    tried = 1;
    esh_fp = open_tap("esh", "esh");
  }
  esh_rec = esh_fp && esh_count < tap_max() && eligible && h->hbe_flag;
  if (!esh_rec) return;
  g_hbe_called = 0;
  esh_bad = esh_state(&esh_in, d, f, t);
}
static void esh_post(ia_sbr_dec_struct *d, ia_sbr_frame_info_data_struct *f, ia_sbr_tables_struct *t) {
  if (!esh_rec) return;
  esh_bad |= esh_state(&esh_out, d, f, t);
  int32_t head[16];
  memcpy(head, esd_head, sizeof(head));
  head[0] = 0x31485345;
  if (esh_bad || g_hbe_called != 1) head[1] = -101;
  fwrite(head, 4, 16, esh_fp);
  fwrite(esd_tin, 4, 1024, esh_fp);
  fwrite(&esh_in, sizeof(esh_in), 1, esh_fp);
  fwrite(g_ehf_par, 4, XO_EHF_PAR_WORDS, esh_fp);
  fwrite(g_eec_ipar_in, 4, XO_EEC_IPAR_WORDS, esh_fp);
  fwrite(g_eec_fpar, 4, XO_EEC_FPAR_WORDS, esh_fp);
  fwrite(g_hbe_cfg, 4, XO_HBE_CFG_WORDS, esh_fp);
  fwrite(d->time_sample_buf, 4, 2048, esh_fp);
  fwrite(&esh_out, sizeof(esh_out), 1, esh_fp);
  fwrite(g_eec_ipar_out, 4, XO_EEC_IPAR_WORDS, esh_fp);
  fflush(esh_fp);
  esh_count++;
}

void esbr_stage_tap_post(ia_sbr_dec_struct *d, ia_sbr_frame_info_data_struct *f, ia_sbr_tables_struct *t, int ret) {
  if (esh_rec) {
    esd_head[1] = ret;
    esd_head[14] = g_ehf_called;
    esd_head[15] = g_eec_called;
    esh_post(d, f, t);
  }
  if (!esd_rec) return;
  esd_head[1] = ret;
  esd_head[14] = g_ehf_called;
  esd_head[15] = g_eec_called;
  esd_state(&esd_out, d, f, t);
  fwrite(esd_head, 4, 16, esd_fp);
  fwrite(esd_tin, 4, 1024, esd_fp);
  fwrite(&esd_in, sizeof(esd_in), 1, esd_fp);
  fwrite(g_ehf_par, 4, XO_EHF_PAR_WORDS, esd_fp);
  fwrite(g_eec_ipar_in, 4, XO_EEC_IPAR_WORDS, esd_fp);
  fwrite(g_eec_fpar, 4, XO_EEC_FPAR_WORDS, esd_fp);
  fwrite(d->time_sample_buf, 4, 2048, esd_fp);
  fwrite(&esd_out, sizeof(esd_out), 1, esd_fp);
  fwrite(g_eec_ipar_out, 4, XO_EEC_IPAR_WORDS, esd_fp);
  fflush(esd_fp);
  esd_count++;
}
