/*
 * libxaac_b200/dropin/ixheaacd_b200_glue.c — the reference-side half of the drop-in.
 *
 * Stage overrides with the EXACT signatures of the reference's stage functions, installed at link time with
 *   -Wl,--wrap=ixheaacd_imdct_process  -Wl,--wrap=ixheaacd_sbr_dec  -Wl,--wrap=ixheaacd_fd_frm_dec
 * (mechanism 2): the reference's own bitstream parser / API layer (L4) keeps calling the names it always
 * called and the calls land on the B200 kernels through the C-ABI of include/xaac_b200.h.
 *   ixheaacd_imdct_process   decoder/ixheaacd_lpfuncs.c:347-353   -> xaac_b200_imdct_process_dev
 *   ixheaacd_sbr_dec         decoder/ixheaacd_sbr_dec.h:219-229   -> xaac_b200_sbr_dec_hq_dev (HQ, with PS) / xaac_b200_sbr_dec_lp_dev
 *   ixheaacd_fd_frm_dec      decoder/ixheaacd_imdct.c:596         -> xaac_b200_usac_fd_frm_dec_dev
 * Frames outside the kernels' subset (LD / ELD object types, 960-sample frames, LPD / FAC transitions, the float eSBR branch,
 * DRC inside the QMF bank ...) go to the reference's own code (__real_*) and are counted; IXHEAACD_B200_STATS=1 prints the
 * counters at exit, IXHEAACD_B200_DISABLE=1 routes everything to the reference.
 *
 * This is the per-call (one decoder instance, one frame at a time) binding: every call ships the channel state to the GPU
 * and back, so that the reference's structs stay authoritative between calls.  It exists to prove the boundary — bit-identical
 * PCM for whole files through the reference's own parser — not for speed; a throughput deployment keeps the state resident and
 * batches streams through the same entry points (bench.py).
 * No CUDA headers are needed here: device memory is handled through xaac_b200_dev_alloc / _h2d / _d2h.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "ixheaacd_b200_pack.h"
#include "ixheaacd_interface.h"
#include "ixheaacd_tns_usac.h"
#include "ixheaacd_acelp_info.h"
#include "ixheaacd_td_mdct.h"
#include "ixheaacd_info.h"
#include "ixheaacd_main.h"
#include "ixheaacd_windows.h"

/* ---- the reference's own implementations (ld --wrap) ---- */
VOID __real_ixheaacd_imdct_process(ia_aac_dec_overlap_info *, WORD32 *, ia_ics_info_struct *, VOID *, const WORD16, WORD32 *,
                                   ia_aac_dec_tables_struct *, WORD32, WORD32, WORD);
WORD32 __real_ixheaacd_sbr_dec(ia_sbr_dec_struct *, WORD16 *, ia_sbr_header_data_struct *, ia_sbr_frame_info_data_struct *,
                               ia_sbr_prev_frame_data_struct *, ia_ps_dec_struct *, ia_sbr_qmf_filter_bank_struct *,
                               ia_sbr_scale_fact_struct *, FLAG, FLAG, WORD32 *, ia_sbr_tables_struct *, ixheaacd_misc_tables *,
                               WORD, ia_pvc_data_struct *, FLAG, WORD32[][64], WORD32, WORD32, VOID *, WORD32, WORD32);
WORD32 __real_ixheaacd_fd_frm_dec(ia_usac_data_struct *usac_data, WORD32 i_ch);

extern const WORD32 ixheaacd_twiddle_table_fft_32x32[514];
extern const WORD32 ixheaacd_pre_post_twid_cos_512[512];
extern const WORD32 ixheaacd_pre_post_twid_sin_512[512];
extern const WORD32 ixheaacd_pre_post_twid_cos_64[64];
extern const WORD32 ixheaacd_pre_post_twid_sin_64[64];

/* ---- one process-wide context (the reference decoder is single-threaded, SURVEY 8b "Threading") ---- */
static struct {
  xaac_b200_ctx *ctx;
  int tried, disabled, stats;
  int have_imdct_rom, have_sbr_rom, have_usac_rom;
  /* device staging for one unit */
  int32_t *d_spec, *d_ovl, *d_out32, *d_err, *d_uovl;
  uint8_t *d_ws, *d_ics;
  int8_t *d_adj;
  int16_t *d_side, *d_tin, *d_pcm;
  xaac_b200_sbr_state *st_hq, *st_ps, *st_lp;
  long n_imdct, n_imdct_ref, n_sbr_hq, n_sbr_ps, n_sbr_lp, n_sbr_ref, n_fd, n_fd_ref;
} G;

static void b200_report(void) {
  if (G.stats)
    fprintf(stderr,
            "[ixheaacd_b200] imdct_process: %ld on the GPU, %ld by the reference; sbr_dec: %ld HQ + %ld HQ/PS + %ld LP on the GPU, "
            "%ld by the reference; fd_frm_dec: %ld on the GPU, %ld by the reference\n",
            G.n_imdct, G.n_imdct_ref, G.n_sbr_hq, G.n_sbr_ps, G.n_sbr_lp, G.n_sbr_ref, G.n_fd, G.n_fd_ref);
  if (G.ctx) xaac_b200_destroy(G.ctx);
  G.ctx = NULL;
}
static void b200_fatal(const char *what) {
  fprintf(stderr, "[ixheaacd_b200] %s: %s\n", what, G.ctx ? xaac_b200_last_error(G.ctx) : "no context");
  exit(3); /* no CPU fallback for a failing device: fail loudly */
}
#define B200(call, what)                 \
  do {                                   \
    if ((call) != XAAC_B200_OK) b200_fatal(what); \
  } while (0)

static xaac_b200_ctx *b200_ctx(void) {
  if (!G.tried) {
    const char *e = getenv("IXHEAACD_B200_DISABLE"), *s = getenv("IXHEAACD_B200_STATS"), *d = getenv("IXHEAACD_B200_DEVICE");
    G.tried = 1;
    G.disabled = e && *e && *e != '0';
    G.stats = s && *s && *s != '0';
    atexit(b200_report);
    if (!G.disabled) {
      if (xaac_b200_create(&G.ctx, d ? atoi(d) : 0) != XAAC_B200_OK) {
        fprintf(stderr, "[ixheaacd_b200] no CUDA device / context (set IXHEAACD_B200_DISABLE=1 to run the reference's own code)\n");
        exit(3);
      }
      B200(xaac_b200_dev_alloc(G.ctx, 4096, (void **)&G.d_spec), "alloc");
      B200(xaac_b200_dev_alloc(G.ctx, 2048, (void **)&G.d_ovl), "alloc");
      B200(xaac_b200_dev_alloc(G.ctx, 4096, (void **)&G.d_uovl), "alloc");
      B200(xaac_b200_dev_alloc(G.ctx, 4096, (void **)&G.d_out32), "alloc");
      B200(xaac_b200_dev_alloc(G.ctx, 16, (void **)&G.d_err), "alloc");
      B200(xaac_b200_dev_alloc(G.ctx, 16, (void **)&G.d_ws), "alloc");
      B200(xaac_b200_dev_alloc(G.ctx, 16, (void **)&G.d_ics), "alloc");
      B200(xaac_b200_dev_alloc(G.ctx, 16, (void **)&G.d_adj), "alloc");
      B200(xaac_b200_dev_alloc(G.ctx, 2 * XAAC_SIDE_WORDS, (void **)&G.d_side), "alloc");
      B200(xaac_b200_dev_alloc(G.ctx, 2048, (void **)&G.d_tin), "alloc");
      B200(xaac_b200_dev_alloc(G.ctx, 2 * 2048 * 2, (void **)&G.d_pcm), "alloc");
    }
  }
  return G.disabled ? NULL : G.ctx;
}

/* ================================ ixheaacd_imdct_process ================================ */
VOID __wrap_ixheaacd_imdct_process(ia_aac_dec_overlap_info *ptr_aac_dec_overlap_info, WORD32 *ptr_spec_coeff,
                                   ia_ics_info_struct *ptr_ics_info, VOID *out_samples, const WORD16 ch_fac, WORD32 *scratch,
                                   ia_aac_dec_tables_struct *ptr_aac_tables, WORD32 object_type, WORD32 ld_mps_present,
                                   WORD slot_element) {
  xaac_b200_ctx *c = b200_ctx();
  if (!c || ptr_ics_info->frame_length != 1024 || object_type == AOT_ER_AAC_LD || object_type == AOT_ER_AAC_ELD ||
      ptr_ics_info->window_sequence > 3) {
    G.n_imdct_ref++;
    __real_ixheaacd_imdct_process(ptr_aac_dec_overlap_info, ptr_spec_coeff, ptr_ics_info, out_samples, ch_fac, scratch,
                                  ptr_aac_tables, object_type, ld_mps_present, slot_element);
    return;
  }
  if (!G.have_imdct_rom) { /* the host passes its own tables, as the reference does to every hot function */
    B200(xaac_b200_set_imdct_rom(c, ptr_aac_tables->pstr_imdct_tables, 7500), "set_imdct_rom");
    G.have_imdct_rom = 1;
  }
  uint8_t ws[2] = {(uint8_t)ptr_aac_dec_overlap_info->window_shape, (uint8_t)ptr_aac_dec_overlap_info->window_sequence};
  uint8_t ics[2] = {(uint8_t)ptr_ics_info->window_sequence, (uint8_t)ptr_ics_info->window_shape};
  int8_t adj = 0;
  static int32_t out[1024];
  B200(xaac_b200_h2d(c, G.d_spec, ptr_spec_coeff, 4096), "h2d spec");
  B200(xaac_b200_h2d(c, G.d_ovl, ptr_aac_dec_overlap_info->ptr_overlap_buf, 2048), "h2d overlap");
  B200(xaac_b200_h2d(c, G.d_ws, ws, 2), "h2d wstate");
  B200(xaac_b200_h2d(c, G.d_ics, ics, 2), "h2d ics");
  B200(xaac_b200_imdct_process_dev(c, G.d_spec, G.d_ovl, G.d_ws, G.d_ics, G.d_out32, G.d_adj, 1, 1, NULL), "imdct_process_dev");
  B200(xaac_b200_d2h(c, out, G.d_out32, 4096), "d2h out");
  B200(xaac_b200_d2h(c, ptr_aac_dec_overlap_info->ptr_overlap_buf, G.d_ovl, 2048), "d2h overlap");
  B200(xaac_b200_d2h(c, ws, G.d_ws, 2), "d2h wstate");
  B200(xaac_b200_d2h(c, &adj, G.d_adj, 1), "d2h qshift_adj");
  WORD32 *po = (WORD32 *)out_samples;
  for (int i = 0; i < 1024; i++) po[ch_fac * i] = out[i];
  ptr_aac_dec_overlap_info->window_shape = ws[0];
  ptr_aac_dec_overlap_info->window_sequence = ws[1];
  ptr_ics_info->qshift_adj = adj;
  G.n_imdct++;
}

/* ================================ ixheaacd_sbr_dec (fixed-point branch) ================================ */
/* record -> the live reference structs: the inverse of pack_sbr_state_lp / pack_ps_state (ixheaacd_b200_pack.h) */
static void unpack_sbr_state_into(const int16_t *st, ia_sbr_dec_struct *d, ia_sbr_prev_frame_data_struct *pv, int low_pow) {
  ia_sbr_qmf_filter_bank_struct *a = &d->str_codec_qmf_bank, *s = &d->str_synthesis_qmf_bank;
  memcpy(a->anal_filter_states, st + XAAC_SBR_ST_ANAL_STATES, 320 * sizeof(int16_t));
  a->core_samples_buffer = a->anal_filter_states + st[XAAC_SBR_ST_ANAL_POS];
  a->filter_pos = (WORD16 *)a->analy_win_coeff + st[XAAC_SBR_ST_ANAL_POS + 1];
  s->ixheaacd_drc_offset = st[XAAC_SBR_ST_SYN_POS];
  s->filter_pos_syn = (WORD16 *)s->p_filter + st[XAAC_SBR_ST_SYN_POS + 1];
  unpack_sf(st + XAAC_SBR_ST_SF, &d->str_sbr_scale_fact);
  const int16_t *misc = st + XAAC_SBR_ST_MISC;
  pv->max_qmf_subband_aac = misc[XAAC_SBR_MISC_MAX_QMF_PREV];
  pv->end_position = misc[XAAC_SBR_MISC_END_POS_PREV];
  for (int i = 0; i < MAX_NUM_NOISE_VALUES; i++) pv->sbr_invf_mode[i] = misc[XAAC_SBR_MISC_INVF_PREV + i];
  a->usb = misc[XAAC_SBR_MISC_CODEC_USB];
  s->lsb = misc[XAAC_SBR_MISC_SYN_LSB];
  s->usb = misc[XAAC_SBR_MISC_SYN_USB];
  unpack_env_state(st + XAAC_SBR_ST_ENV, &d->str_sbr_calc_env);
  memcpy(s->filter_states, st + XAAC_SBR_ST_SYN_STATES, 1280 * sizeof(int16_t));
  memcpy(d->str_hf_generator.bw_array_prev, st + XAAC_SBR_ST_BW_PREV, 6 * sizeof(int32_t));
  const int32_t *lpc = (const int32_t *)(st + XAAC_SBR_ST_LPC);
  for (int i = 0; i < 2; i++) {
    memcpy(d->str_hf_generator.lpc_filt_states_real[i], lpc + 128 * i, 32 * sizeof(int32_t));
    if (!low_pow && d->str_hf_generator.lpc_filt_states_imag[i])
      memcpy(d->str_hf_generator.lpc_filt_states_imag[i], lpc + 128 * i + 64, 32 * sizeof(int32_t));
  }
  memcpy(d->ptr_sbr_overlap_buf, st + XAAC_SBR_ST_OV, (low_pow ? 6 * 64 : 6 * 128) * sizeof(int32_t));
}
static void unpack_ps_state_into(const int16_t *p, ia_ps_dec_struct *ps, ia_sbr_qmf_filter_bank_struct *bank_r,
                                 ia_sbr_scale_fact_struct *sf_r) {
  memcpy(ps->delay_buf_qmf_ap_re_im, p + XAAC_PS_ST_AP, 128 * sizeof(int16_t));
  memcpy(ps->delay_buf_qmf_ld_re_im, p + XAAC_PS_ST_LD, 336 * sizeof(int16_t));
  memcpy(ps->delay_buf_qmf_sd_re_im, p + XAAC_PS_ST_SD, 58 * sizeof(int16_t));
  memcpy(ps->delay_buf_qmf_ser_re_im, p + XAAC_PS_ST_SER, 960 * sizeof(int16_t));
  memcpy(ps->delay_buf_qmf_sub_re_im, p + XAAC_PS_ST_SUB, 64 * sizeof(int16_t));
  memcpy(ps->delay_buf_qmf_sub_ser_re_im, p + XAAC_PS_ST_SUB_SER, 480 * sizeof(int16_t));
  const int16_t *hv = p + XAAC_PS_ST_HVEC;
  memcpy(ps->h11_h12_vec, hv, 96); memcpy(ps->h21_h22_vec, hv + 48, 96); memcpy(ps->H11_H12, hv + 96, 96);
  memcpy(ps->H21_H22, hv + 144, 96); memcpy(ps->delta_h11_h12, hv + 192, 96); memcpy(ps->delta_h21_h22, hv + 240, 96);
  const int16_t *idx = p + XAAC_PS_ST_IDX;
  for (int i = 0; i < 3; i++) ps->delay_buf_idx_ser[i] = idx[XAAC_PS_IDX_SER + i];
  ps->delay_buf_idx = idx[XAAC_PS_IDX_DELAY];
  ps->delay_buf_idx_long = idx[XAAC_PS_IDX_DELAY_LONG];
  ps->delay_buffer_scale = idx[XAAC_PS_IDX_SCALE];
  ps->usb = idx[XAAC_PS_IDX_USB];
  bank_r->lsb = idx[XAAC_PS_IDX_LSB_R];
  bank_r->usb = idx[XAAC_PS_IDX_USB_R];
  const int32_t *pk = (const int32_t *)(p + XAAC_PS_ST_PEAK);
  memcpy(ps->peak_decay_diff, pk, 80); memcpy(ps->energy_prev, pk + 20, 80); memcpy(ps->peak_decay_diff_prev, pk + 40, 80);
  const int32_t *hy = (const int32_t *)(p + XAAC_PS_ST_HYB);
  for (int b = 0; b < 3; b++) {
    memcpy(ps->str_hybrid.ptr_qmf_buf_re[b], hy + 24 * b, 48);
    memcpy(ps->str_hybrid.ptr_qmf_buf_im[b], hy + 24 * b + 12, 48);
  }
  memcpy(bank_r->filter_states, p + XAAC_PS_ST_SYN_STATES_R, 1280 * sizeof(int16_t));
  bank_r->ixheaacd_drc_offset = p[XAAC_PS_ST_SYN_POS_R];
  bank_r->filter_pos_syn = (WORD16 *)bank_r->p_filter + p[XAAC_PS_ST_SYN_POS_R + 1];
  unpack_sf(p + XAAC_PS_ST_SF_R, sf_r);
}

WORD32 __wrap_ixheaacd_sbr_dec(ia_sbr_dec_struct *ptr_sbr_dec, WORD16 *ptr_time_data, ia_sbr_header_data_struct *ptr_header_data,
                               ia_sbr_frame_info_data_struct *ptr_frame_data, ia_sbr_prev_frame_data_struct *ptr_frame_data_prev,
                               ia_ps_dec_struct *ptr_ps_dec, ia_sbr_qmf_filter_bank_struct *ptr_qmf_synth_bank_r,
                               ia_sbr_scale_fact_struct *ptr_sbr_sf_r, FLAG apply_processing, FLAG low_pow_flag,
                               WORD32 *ptr_work_buf_core, ia_sbr_tables_struct *sbr_tables_ptr,
                               ixheaacd_misc_tables *pstr_common_tables, WORD ch_fac, ia_pvc_data_struct *ptr_pvc_data_str,
                               FLAG drc_on, WORD32 drc_sbr_factors[][64], WORD32 audio_object_type, WORD32 ldmps_present,
                               VOID *self, WORD32 heaac_mps_present, WORD32 ec_flag) {
  xaac_b200_ctx *c = b200_ctx();
  /* the low-power branch never runs PS (the caller may still hand over the PS instance of the element) */
  const int ps_present = !low_pow_flag && ptr_ps_dec != NULL && ptr_qmf_synth_bank_r != NULL && ptr_sbr_sf_r != NULL;
  const int eligible = c && !ptr_header_data->enh_sbr && ptr_header_data->num_time_slots == 16 && ptr_header_data->time_step == 2 &&
                       audio_object_type != AOT_ER_AAC_ELD && audio_object_type != AOT_ER_AAC_LD && !ldmps_present && !drc_on &&
                       !heaac_mps_present && !ec_flag && ptr_sbr_dec->str_codec_qmf_bank.no_channels == 32 &&
                       ptr_sbr_dec->str_synthesis_qmf_bank.no_channels == 64;
  if (!eligible) {
    G.n_sbr_ref++;
    return __real_ixheaacd_sbr_dec(ptr_sbr_dec, ptr_time_data, ptr_header_data, ptr_frame_data, ptr_frame_data_prev, ptr_ps_dec,
                                   ptr_qmf_synth_bank_r, ptr_sbr_sf_r, apply_processing, low_pow_flag, ptr_work_buf_core,
                                   sbr_tables_ptr, pstr_common_tables, ch_fac, ptr_pvc_data_str, drc_on, drc_sbr_factors,
                                   audio_object_type, ldmps_present, self, heaac_mps_present, ec_flag);
  }
  if (!G.have_sbr_rom) {
    B200(xaac_b200_set_qmf_rom(c, sbr_tables_ptr->qmf_dec_tables_ptr, 3464), "set_qmf_rom");
    B200(xaac_b200_set_env_rom(c, sbr_tables_ptr->env_calc_tables_ptr, 2404, pstr_common_tables, 2470), "set_env_rom");
    B200(xaac_b200_set_ps_rom(c, sbr_tables_ptr->ps_tables_ptr, 1230), "set_ps_rom");
    G.have_sbr_rom = 1;
  }
  static int16_t side[XAAC_SIDE_WORDS], st[XAAC_SBR_ST_WORDS], pst[XAAC_PS_ST_WORDS], tin[1024], pcm[2 * 2048];
  int32_t err = 0;
  xaac_b200_sbr_state **slot = low_pow_flag ? &G.st_lp : (ps_present ? &G.st_ps : &G.st_hq);
  if (!*slot)
    B200(xaac_b200_sbr_state_create(c, 1, low_pow_flag ? XAAC_B200_SBR_STATE_LP : (ps_present ? 1 : 0), slot), "sbr_state_create");
  pack_side(side, ptr_sbr_dec, ptr_header_data, ptr_frame_data, ptr_frame_data_prev, ps_present ? ptr_ps_dec : NULL,
            apply_processing);
  pack_sbr_state_lp(st, ptr_sbr_dec, ptr_frame_data_prev, low_pow_flag);
  if (ps_present) pack_ps_state(pst, ptr_ps_dec, ptr_qmf_synth_bank_r, ptr_sbr_sf_r);
  for (int i = 0; i < 1024; i++) tin[i] = ptr_time_data[ch_fac * i];
  B200(xaac_b200_sbr_state_upload(c, *slot, st, ps_present ? pst : NULL), "sbr_state_upload");
  B200(xaac_b200_h2d(c, G.d_side, side, sizeof(side)), "h2d side");
  B200(xaac_b200_h2d(c, G.d_tin, tin, sizeof(tin)), "h2d time");
  if (low_pow_flag)
    B200(xaac_b200_sbr_dec_lp_dev(c, *slot, G.d_side, G.d_tin, G.d_pcm, 1, G.d_err, NULL), "sbr_dec_lp_dev");
  else
    B200(xaac_b200_sbr_dec_hq_dev(c, *slot, G.d_side, G.d_tin, G.d_pcm, G.d_err, NULL), "sbr_dec_hq_dev");
  B200(xaac_b200_d2h(c, &err, G.d_err, 4), "d2h err");
  B200(xaac_b200_d2h(c, pcm, G.d_pcm, ps_present ? 8192 : 4096), "d2h pcm");
  B200(xaac_b200_sbr_state_download(c, *slot, st, ps_present ? pst : NULL), "sbr_state_download");
  unpack_sbr_state_into(st, ptr_sbr_dec, ptr_frame_data_prev, low_pow_flag);
  if (ps_present) unpack_ps_state_into(pst, ptr_ps_dec, ptr_qmf_synth_bank_r, ptr_sbr_sf_r);
  if (err == 0) {
    const int run_ps = ps_present && side[XAAC_SIDE_PS];
    if (ps_present) {
      for (int i = 0; i < 2048; i++) {
        ptr_time_data[ch_fac * i] = pcm[2 * i];
        if (run_ps) ptr_time_data[ch_fac * i + 1] = pcm[2 * i + 1];
      }
    } else {
      for (int i = 0; i < 2048; i++) ptr_time_data[ch_fac * i] = pcm[i];
    }
  }
  if (low_pow_flag) G.n_sbr_lp++; else if (ps_present && side[XAAC_SIDE_PS]) G.n_sbr_ps++; else G.n_sbr_hq++;
  return err;
}

/* ================================ ixheaacd_fd_frm_dec ================================ */
WORD32 __wrap_ixheaacd_fd_frm_dec(ia_usac_data_struct *usac_data, WORD32 i_ch) {
  xaac_b200_ctx *c = b200_ctx();
  const int seq = usac_data->window_sequence[i_ch];
  if (!c || usac_data->ccfl != 1024 || usac_data->ec_flag || usac_data->td_frame_prev[i_ch] || usac_data->fac_data_present[i_ch] ||
      seq < 0 || seq > 4) {
    G.n_fd_ref++;
    return __real_ixheaacd_fd_frm_dec(usac_data, i_ch);
  }
  if (!G.have_usac_rom) { /* XAAC_UROM_* blob from the reference's global tables */
    static int32_t blob[XAAC_UROM_BYTES / 4];
    int32_t *p = blob;
    memcpy(p, ixheaacd_twiddle_table_fft_32x32, 514 * 4); p += 514;
    memcpy(p, ixheaacd_pre_post_twid_cos_512, 512 * 4); p += 512;
    memcpy(p, ixheaacd_pre_post_twid_sin_512, 512 * 4); p += 512;
    memcpy(p, ixheaacd_pre_post_twid_cos_64, 64 * 4); p += 64;
    memcpy(p, ixheaacd_pre_post_twid_sin_64, 64 * 4); p += 64;
    memcpy(p, ixheaacd_sine_win_1024, 1024 * 4); p += 1024;
    memcpy(p, ixheaacd_kbd_win1024, 1024 * 4); p += 1024;
    memcpy(p, ixheaacd_sine_win_128, 128 * 4); p += 128;
    memcpy(p, ixheaacd_kbd_win128, 128 * 4);
    B200(xaac_b200_set_usac_rom(c, blob, sizeof(blob)), "set_usac_rom");
    G.have_usac_rom = 1;
  }
  uint8_t ws = (uint8_t)usac_data->window_shape_prev[i_ch];
  uint8_t ics[2] = {(uint8_t)seq, (uint8_t)usac_data->window_shape[i_ch]};
  B200(xaac_b200_h2d(c, G.d_spec, usac_data->coef_fix[i_ch], 4096), "h2d coef");
  B200(xaac_b200_h2d(c, G.d_uovl, usac_data->overlap_data_ptr[i_ch], 4096), "h2d overlap");
  B200(xaac_b200_h2d(c, G.d_ws, &ws, 1), "h2d wstate");
  B200(xaac_b200_h2d(c, G.d_ics, ics, 2), "h2d ics");
  B200(xaac_b200_usac_fd_frm_dec_dev(c, G.d_spec, G.d_uovl, G.d_ws, G.d_ics, G.d_out32, 1, NULL), "usac_fd_frm_dec_dev");
  B200(xaac_b200_d2h(c, usac_data->output_data_ptr[i_ch], G.d_out32, 4096), "d2h out");
  B200(xaac_b200_d2h(c, usac_data->overlap_data_ptr[i_ch], G.d_uovl, 4096), "d2h overlap");
  G.n_fd++;
  return 0;
}
