/*
 * oracle/ref_taps.c — TEST INFRASTRUCTURE ONLY.
 *
 * Stage taps for the UNMODIFIED reference decoder, installed at link time with `ld --wrap`:
 * the reference testbench (test/decoder/*.c) + libxaacdec.a + this file give oracle/_ref/xaacdec_tap, which
 * decodes a real bitstream exactly like xaacdec and, when XAAC_TAP_FILE is set, appends one binary record
 * per stage call (inputs, state-before, outputs, state-after).  tools/make_golden.py turns those records
 * into the small fixtures under tests/golden/.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#define REF_SHIM_HEADERS_ONLY
#include "ref_headers.h"

static FILE *tap_fp(void) {
  static FILE *fp = NULL;
  static int tried = 0;
  if (!tried) {
    const char *p = getenv("XAAC_TAP_FILE");
    tried = 1;
    if (p && *p) fp = fopen(p, "wb");
  }
  return fp;
}
/* XAAC_TAP_STAGES: comma-separated subset of {imd,hfg,env,sbr,ps}; default = all */
static int tap_on(const char *stage) {
  const char *p = getenv("XAAC_TAP_STAGES");
  if (!p || !*p) return 1;
  return strstr(p, stage) != NULL;
}
static int tap_limit(void) {
  static int lim = -1;
  if (lim < 0) {
    const char *p = getenv("XAAC_TAP_MAX");
    lim = p ? atoi(p) : 1000000;
  }
  return lim;
}

/* ---- ixheaacd_imdct_process (decoder/ixheaacd_lpfuncs.c:347) -------------------------------------------
 * record: int32 magic 'IMD1', int32 hdr[8] = {frame_length, object_type, ch_fac, prev_shape, prev_seq,
 *         win_seq, win_shape, qshift_adj}, int32 spec[1024], ovl_before[512], out[1024], ovl_after[512] */
VOID __real_ixheaacd_imdct_process(ia_aac_dec_overlap_info *, WORD32 *, ia_ics_info_struct *, VOID *,
                                   const WORD16, WORD32 *, ia_aac_dec_tables_struct *, WORD32, WORD32, WORD);

VOID __wrap_ixheaacd_imdct_process(ia_aac_dec_overlap_info *ovl, WORD32 *spec, ia_ics_info_struct *ics,
                                   VOID *out, const WORD16 ch_fac, WORD32 *scratch,
                                   ia_aac_dec_tables_struct *tabs, WORD32 object_type, WORD32 ld_mps,
                                   WORD slot) {
  static int count = 0;
  FILE *fp = tap_fp();
  int rec = fp && tap_on("imd") && ics->frame_length == 1024 && count < tap_limit();
  int32_t hdr[9];
  int32_t spec_in[1024], ovl_in[512];
  if (rec) {
    hdr[0] = 0x31444d49;
    hdr[1] = ics->frame_length;
    hdr[2] = object_type;
    hdr[3] = ch_fac;
    hdr[4] = ovl->window_shape;
    hdr[5] = ovl->window_sequence;
    hdr[6] = ics->window_sequence;
    hdr[7] = ics->window_shape;
    memcpy(spec_in, spec, sizeof(spec_in));
    memcpy(ovl_in, ovl->ptr_overlap_buf, sizeof(ovl_in));
  }
  __real_ixheaacd_imdct_process(ovl, spec, ics, out, ch_fac, scratch, tabs, object_type, ld_mps, slot);
  if (rec) {
    int32_t o[1024];
    const WORD32 *po = (const WORD32 *)out;
    for (int i = 0; i < 1024; i++) o[i] = po[ch_fac * i];
    hdr[8] = ics->qshift_adj;
    fwrite(hdr, 4, 9, fp);
    fwrite(spec_in, 4, 1024, fp);
    fwrite(ovl_in, 4, 512, fp);
    fwrite(o, 4, 1024, fp);
    fwrite(ovl->ptr_overlap_buf, 4, 512, fp);
    fflush(fp);
    count++;
  }
}

/* ---- ixheaacd_hf_generator (decoder/ixheaacd_lpp_tran.c:956), HQ path --------------------------------------
 * record: int32 magic 'HFG1', int16 prm[80] (layout XO_HF_*), int32 bw_prev_in[6], int32 lpc[2][128],
 *         int32 matrix_in[38][128], int32 matrix_out[38][128], int32 bw_prev_out[6], int32 hb_scale */
VOID __real_ixheaacd_hf_generator(ia_sbr_hf_generator_struct *, ia_sbr_scale_fact_struct *, WORD32 **, WORD32 **,
                                  WORD32, WORD32, WORD32, WORD32, WORD32, WORD32 *, WORD32 *, WORD32 *, WORD);

VOID __wrap_ixheaacd_hf_generator(ia_sbr_hf_generator_struct *hf, ia_sbr_scale_fact_struct *sf, WORD32 **re,
                                  WORD32 **im, WORD32 factor, WORD32 start_idx, WORD32 stop_idx, WORD32 num_if_bands,
                                  WORD32 max_qmf_subband, WORD32 *invf, WORD32 *invf_prev, WORD32 *sub_sig_x,
                                  WORD aot) {
  static int count = 0;
  FILE *fp = tap_fp();
  ia_transposer_settings_struct *set = hf->pstr_settings;
  int rec = fp && tap_on("hfg") && count < tap_limit() && set->num_columns == 32;
  int16_t prm[80];
  int32_t bw_in[6], lpc[2][128];
  static int32_t m_in[38][128];
  if (rec) {
    memset(prm, 0, sizeof(prm));
    prm[0] = set->num_patches; prm[1] = set->start_patch; prm[2] = set->stop_patch; prm[3] = set->num_columns;
    for (int i = 0; i < MAX_NUM_NOISE_VALUES; i++) prm[4 + i] = set->bw_borders[i];
    for (int p = 0; p < MAX_NUM_PATCHES; p++) {
      int16_t *q = prm + 14 + 6 * p;
      q[0] = set->str_patch_param[p].src_start_band; q[1] = set->str_patch_param[p].src_end_band;
      q[2] = set->str_patch_param[p].guard_start_band; q[3] = set->str_patch_param[p].dst_start_band;
      q[4] = set->str_patch_param[p].dst_end_band; q[5] = set->str_patch_param[p].num_bands_in_patch;
    }
    prm[50] = (int16_t)factor; prm[51] = (int16_t)num_if_bands; prm[52] = (int16_t)start_idx; prm[53] = (int16_t)stop_idx;
    for (int i = 0; i < MAX_NUM_NOISE_VALUES; i++) { prm[54 + i] = (int16_t)invf[i]; prm[64 + i] = (int16_t)invf_prev[i]; }
    prm[74] = sf->ov_lb_scale; prm[75] = sf->lb_scale; prm[76] = (int16_t)max_qmf_subband;
    for (int i = 0; i < 6; i++) bw_in[i] = hf->bw_array_prev[i];
    for (int i = 0; i < 2; i++) {
      memcpy(lpc[i], hf->lpc_filt_states_real[i], 64 * 4);
      memcpy(lpc[i] + 64, hf->lpc_filt_states_imag[i], 64 * 4);
    }
    for (int i = 0; i < 38; i++) { memcpy(m_in[i], re[i], 64 * 4); memcpy(m_in[i] + 64, im[i], 64 * 4); }
  }
  __real_ixheaacd_hf_generator(hf, sf, re, im, factor, start_idx, stop_idx, num_if_bands, max_qmf_subband, invf,
                               invf_prev, sub_sig_x, aot);
  if (rec) {
    int32_t magic = 0x31474648, hb = sf->hb_scale;
    fwrite(&magic, 4, 1, fp);
    fwrite(prm, 2, 80, fp);
    fwrite(bw_in, 4, 6, fp);
    fwrite(lpc, 4, 256, fp);
    fwrite(m_in, 4, 38 * 128, fp);
    for (int i = 0; i < 38; i++) { fwrite(re[i], 4, 64, fp); fwrite(im[i], 4, 64, fp); }
    fwrite(hf->bw_array_prev, 4, 6, fp);
    fwrite(&hb, 4, 1, fp);
    fflush(fp);
    count++;
  }
}

/* ---- ixheaacd_calc_sbrenvelope (decoder/ixheaacd_env_calc.c:692), HQ path ------------------------------------
 * record: int32 magic 'ENV1', int16 prm[656] (XO_ENV_*), int16 sf_in[8], int16 st_in[232], int32 m_in[38][128],
 *         int32 m_out[38][128], int16 sf_out[8], int16 st_out[232], int32 err */
#include "ref_pack.h"
IA_ERRORCODE __real_ixheaacd_calc_sbrenvelope(ia_sbr_scale_fact_struct *, ia_sbr_calc_env_struct *,
                                              ia_sbr_header_data_struct *, ia_sbr_frame_info_data_struct *,
                                              ia_sbr_prev_frame_data_struct *, WORD32 **, WORD32 **, WORD16 *, FLAG,
                                              ia_sbr_tables_struct *, ixheaacd_misc_tables *, WORD32 *, WORD32);

IA_ERRORCODE __wrap_ixheaacd_calc_sbrenvelope(ia_sbr_scale_fact_struct *sf, ia_sbr_calc_env_struct *ce,
                                              ia_sbr_header_data_struct *h, ia_sbr_frame_info_data_struct *f,
                                              ia_sbr_prev_frame_data_struct *pv, WORD32 **re, WORD32 **im,
                                              WORD16 *deg, FLAG low_pow, ia_sbr_tables_struct *t,
                                              ixheaacd_misc_tables *ct, WORD32 *qmf, WORD32 aot) {
  static int count = 0;
  FILE *fp = tap_fp();
  int rec = fp && tap_on("env") && count < tap_limit() && !low_pow && h->num_time_slots == 16;
  static int16_t prm[XO_ENV_PRM_WORDS], sfr[8], st[XO_ENV_ST_WORDS];
  if (rec) {
    int32_t magic = 0x31564e45;
    pack_env_prm(prm, h, f, pv);
    pack_sf(sfr, sf);
    pack_env_state(st, ce);
    fwrite(&magic, 4, 1, fp);
    fwrite(prm, 2, XO_ENV_PRM_WORDS, fp);
    fwrite(sfr, 2, 8, fp);
    fwrite(st, 2, XO_ENV_ST_WORDS, fp);
    for (int i = 0; i < 38; i++) { fwrite(re[i], 4, 64, fp); fwrite(im[i], 4, 64, fp); }
  }
  IA_ERRORCODE err = __real_ixheaacd_calc_sbrenvelope(sf, ce, h, f, pv, re, im, deg, low_pow, t, ct, qmf, aot);
  if (rec) {
    int32_t e = (int32_t)err;
    for (int i = 0; i < 38; i++) { fwrite(re[i], 4, 64, fp); fwrite(im[i], 4, 64, fp); }
    pack_sf(sfr, sf);
    pack_env_state(st, ce);
    fwrite(sfr, 2, 8, fp);
    fwrite(st, 2, XO_ENV_ST_WORDS, fp);
    fwrite(&e, 4, 1, fp);
    fflush(fp);
    count++;
  }
  return err;
}

/* ---- ixheaacd_sbr_dec (decoder/ixheaacd_sbr_dec.c:662), fixed-point HQ path, whole stage ------------------------
 * record: int32 magic 'SBR1', int32 hdr[7] = {apply_processing, ch_fac, aot, ps_present, ret, 0, 0},
 *   int16 side[XO_SIDE_WORDS], st_in[XO_SBR_ST_WORDS], ps_in[XO_PS_ST_WORDS], time_in[1024],
 *   st_out[XO_SBR_ST_WORDS], ps_out[XO_PS_ST_WORDS], out_l[2048], out_r[2048] */
WORD32 __real_ixheaacd_sbr_dec(ia_sbr_dec_struct *, WORD16 *, ia_sbr_header_data_struct *, ia_sbr_frame_info_data_struct *,
                               ia_sbr_prev_frame_data_struct *, ia_ps_dec_struct *, ia_sbr_qmf_filter_bank_struct *,
                               ia_sbr_scale_fact_struct *, FLAG, FLAG, WORD32 *, ia_sbr_tables_struct *,
                               ixheaacd_misc_tables *, WORD, ia_pvc_data_struct *, FLAG, WORD32[][64], WORD32, WORD32,
                               VOID *, WORD32, WORD32);

void esbr_stage_tap_pre(ia_sbr_dec_struct *d, ia_sbr_header_data_struct *h, ia_sbr_frame_info_data_struct *f,
                        ia_ps_dec_struct *ps, ia_sbr_tables_struct *t, int apply, int low_pow, int aot, int ldmps, int drc_on);
void esbr_stage_tap_post(ia_sbr_dec_struct *d, ia_sbr_frame_info_data_struct *f, ia_sbr_tables_struct *t, int ret);

WORD32 __wrap_ixheaacd_sbr_dec(ia_sbr_dec_struct *d, WORD16 *time, ia_sbr_header_data_struct *h,
                               ia_sbr_frame_info_data_struct *f, ia_sbr_prev_frame_data_struct *pv, ia_ps_dec_struct *ps,
                               ia_sbr_qmf_filter_bank_struct *bank_r, ia_sbr_scale_fact_struct *sf_r, FLAG apply,
                               FLAG low_pow, WORD32 *work, ia_sbr_tables_struct *t, ixheaacd_misc_tables *ct, WORD ch_fac,
                               ia_pvc_data_struct *pvc, FLAG drc_on, WORD32 drc[][64], WORD32 aot, WORD32 ldmps,
                               VOID *self, WORD32 mps, WORD32 ec) {
  static int count = 0;
  FILE *fp = tap_fp();
  int rec = fp && tap_on(low_pow ? "slp" : "sbr") && count < tap_limit() && !h->enh_sbr && h->num_time_slots == 16 &&
            aot != AOT_ER_AAC_ELD && aot != AOT_ER_AAC_LD && !ldmps && !drc_on;
  static int16_t side[XO_SIDE_WORDS], st[XO_SBR_ST_WORDS], pst[XO_PS_ST_WORDS], tin[1024], out[2048];
  int ps_present = 0;
  if (rec) {
    int32_t magic = 0x31524253;
    ps_present = (ps != NULL && bank_r != NULL && sf_r != NULL);
    pack_side(side, d, h, f, pv, ps_present ? ps : NULL, apply);
    pack_sbr_state_lp(st, d, pv, low_pow);
    memset(pst, 0, sizeof(pst));
    if (ps_present) pack_ps_state(pst, ps, bank_r, sf_r);
    for (int i = 0; i < 1024; i++) tin[i] = time[ch_fac * i];
    fwrite(&magic, 4, 1, fp);
  }
  esbr_stage_tap_pre(d, h, f, ps, t, apply, low_pow, aot, ldmps, drc_on);
  WORD32 ret = __real_ixheaacd_sbr_dec(d, time, h, f, pv, ps, bank_r, sf_r, apply, low_pow, work, t, ct, ch_fac, pvc,
                                       drc_on, drc, aot, ldmps, self, mps, ec);
  esbr_stage_tap_post(d, f, t, ret);
  if (rec) {
    int32_t hdr[7] = {apply, ch_fac, aot, ps_present, ret, low_pow, 0};
    fwrite(hdr, 4, 7, fp);
    fwrite(side, 2, XO_SIDE_WORDS, fp);
    fwrite(st, 2, XO_SBR_ST_WORDS, fp);
    fwrite(pst, 2, XO_PS_ST_WORDS, fp);
    fwrite(tin, 2, 1024, fp);
    pack_sbr_state_lp(st, d, pv, low_pow);
    if (ps_present) pack_ps_state(pst, ps, bank_r, sf_r);
    fwrite(st, 2, XO_SBR_ST_WORDS, fp);
    fwrite(pst, 2, XO_PS_ST_WORDS, fp);
    for (int i = 0; i < 2048; i++) out[i] = time[ch_fac * i];
    fwrite(out, 2, 2048, fp);
    for (int i = 0; i < 2048; i++) out[i] = (ch_fac > 1) ? time[ch_fac * i + 1] : 0;
    fwrite(out, 2, 2048, fp);
    fflush(fp);
    count++;
  }
  return ret;
}
