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
#include "ixheaacd_b200_pack_ps_flt.h"
#include "ixheaacd_b200_pack_spec.h"
#include "ixheaacd_b200_pack_sd.h"
#include "ixheaacd_audioobjtypes.h"
#include "ixheaacd_interface.h"
#include "ixheaacd_tns_usac.h"
#include "ixheaacd_acelp_info.h"
#include "ixheaacd_td_mdct.h"
#include "ixheaacd_info.h"
#include "ixheaacd_main.h"
#include "ixheaacd_windows.h"
#include "ixheaacd_sbrqmftrans.h"
#include "ixheaac_esbr_rom.h"

/* ---- the reference's own implementations (ld --wrap) ---- */
VOID __real_ixheaacd_imdct_process(ia_aac_dec_overlap_info *, WORD32 *, ia_ics_info_struct *, VOID *, const WORD16, WORD32 *,
                                   ia_aac_dec_tables_struct *, WORD32, WORD32, WORD);
WORD32 __real_ixheaacd_sbr_dec(ia_sbr_dec_struct *, WORD16 *, ia_sbr_header_data_struct *, ia_sbr_frame_info_data_struct *,
                               ia_sbr_prev_frame_data_struct *, ia_ps_dec_struct *, ia_sbr_qmf_filter_bank_struct *,
                               ia_sbr_scale_fact_struct *, FLAG, FLAG, WORD32 *, ia_sbr_tables_struct *, ixheaacd_misc_tables *,
                               WORD, ia_pvc_data_struct *, FLAG, WORD32[][64], WORD32, WORD32, VOID *, WORD32, WORD32);
WORD32 __real_ixheaacd_fd_frm_dec(ia_usac_data_struct *usac_data, WORD32 i_ch);
IA_ERRORCODE __real_ixheaacd_channel_pair_process(ia_aac_dec_channel_info_struct *[], WORD32, ia_aac_dec_tables_struct *, WORD32, WORD32,
                                                  WORD32, WORD32, WORD32 *, WORD32 *, void *);

extern const FLOAT32 ixheaac_twiddle_table_fft_float[514];
extern const FLOAT32 ixheaac_twidle_tbl_48[64];
extern const FLOAT32 ixheaac_twidle_tbl_24[32];
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
  /* float eSBR stage, one unit */
  int have_esbr_rom;
  float *e_q[6], *e_bw, *e_ec, *e_hbe, *e_fpar, *e_tin, *e_out;
  int32_t *e_anal, *e_apos, *e_synth, *e_spos, *e_patch, *e_hbecfg, *e_hfpar, *e_ipar, *e_rg, *e_err;
  float *e_ps_state, *e_ps_left, *e_ps_right, *e_ps_side, *e_out_r; /* mono + PS element: float parametric stereo */
  int32_t *e_synth_r, *e_spos_r;
  long n_esbr_ps, n_esbr_rebuilt, n_esbr_bypass, n_esbr_tes;
  /* pre-IMDCT spectral stage */
  int have_block_rom;
  uint8_t *d_sps;
  int32_t *d_sps_spec, *d_sps_seed;
  long n_cpp, n_cpp_ref, n_cpp_ms, n_cpp_tns, n_cpp_pns;
  /* SBR side-info dequantisation */
  int16_t *d_sd;
  int have_sd_rom;
  long n_sd, n_sd_ref, n_sd_coupled, n_sd_concealed;
  int16_t *d_psd;
  long n_psd;
  int32_t last_err[6];
  long n_imdct, n_imdct_ref, n_sbr_hq, n_sbr_ps, n_sbr_lp, n_sbr_ref, n_fd, n_fd_ref, n_esbr, n_esbr_hbe, n_esbr_ref;
} G;

static void b200_report(void) {
  if (G.stats)
    fprintf(stderr,
            "[ixheaacd_b200] imdct_process: %ld on the GPU, %ld by the reference; sbr_dec: %ld HQ + %ld HQ/PS + %ld LP on the GPU, "
            "%ld by the reference; fd_frm_dec: %ld on the GPU, %ld by the reference; eSBR sbr_dec: %ld + %ld with HBE + %ld with PS "
            "on the GPU (%ld with the limiter tables rebuilt between the stage halves; %ld with inter-TES) + %ld pass-through, %ld by the reference\n",
            G.n_imdct, G.n_imdct_ref, G.n_sbr_hq, G.n_sbr_ps, G.n_sbr_lp, G.n_sbr_ref, G.n_fd, G.n_fd_ref, G.n_esbr, G.n_esbr_hbe,
            G.n_esbr_ps, G.n_esbr_rebuilt, G.n_esbr_tes, G.n_esbr_bypass, G.n_esbr_ref);
  if (G.stats)
    fprintf(stderr, "[ixheaacd_b200] channel_pair_process: %ld on the GPU (%ld with M/S or intensity bands, %ld with TNS, %ld with PNS), %ld by the "
            "reference\n", G.n_cpp, G.n_cpp_ms, G.n_cpp_tns, G.n_cpp_pns, G.n_cpp_ref);
  if (G.stats)
    fprintf(stderr, "[ixheaacd_b200] dec_sbrdata: %ld on the GPU (%ld coupled pairs, %ld with a concealed channel), %ld by the reference\n",
            G.n_sd, G.n_sd_coupled, G.n_sd_concealed, G.n_sd_ref);
  if (G.stats) fprintf(stderr, "[ixheaacd_b200] decode_ps_data: %ld on the GPU\n", G.n_psd);
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
    G.stats = s && *s && *s != '0' ? atoi(s) : 0;
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

/* ================================ ixheaacd_sbr_dec (float eSBR branch, USAC channels) ================================ */
/* The eSBR branch of ixheaacd_sbr_dec (decoder/ixheaacd_sbr_dec.c:812-1006) for a USAC channel, with or without the harmonic
 * transposer: xaac_b200_esbr_dec_dev / xaac_b200_esbr_dec_hbe_dev.  Frames on which ixheaacd_sbr_env_calc rebuilds its limiter
 * tables (reset_flag, a change of sbr_patching_mode: ixheaacd_createlimiterbands needs THIS frame's patch table, which the HF
 * generator of the same call produces) and everything outside the kernels' subset stay with the reference's code. */
static void esbr_rom_once(xaac_b200_ctx *c, ia_sbr_tables_struct *t) {
  if (G.have_esbr_rom) return;
  ia_qmf_dec_tables_struct *q = t->qmf_dec_tables_ptr;
  static int32_t er[XAAC_EROM_BYTES / 4];
  memcpy((char *)er + XAAC_EROM_QMF_C, q->esbr_qmf_c, 1280 * 4);
  memcpy((char *)er + XAAC_EROM_W32, q->esbr_w_32, 60 * 4);
  memcpy((char *)er + XAAC_EROM_SINCOS_L64, q->esbr_sin_cos_twiddle_l64, 64 * 4);
  memcpy((char *)er + XAAC_EROM_ALTSIN_L64, q->esbr_alt_sin_twiddle_l64, 32 * 4);
  memcpy((char *)er + XAAC_EROM_W16, q->esbr_w_16, 24 * 4);
  memcpy((char *)er + XAAC_EROM_SINCOS_L32, q->esbr_sin_cos_twiddle_l32, 32 * 4);
  memcpy((char *)er + XAAC_EROM_ALTSIN_L32, q->esbr_alt_sin_twiddle_l32, 16 * 4);
  memcpy((char *)er + XAAC_EROM_TCOS_L32, q->esbr_t_cos_sin_l32, 64 * 4);
  B200(xaac_b200_set_esbr_rom(c, er, sizeof(er)), "set_esbr_rom");
  B200(xaac_b200_set_esbr_envcalc_rom(c, ixheaac_random_phase, 4096), "set_esbr_envcalc_rom");
  static float hr[XAAC_HROM_WORDS];
  memcpy(hr + XAAC_HROM_WIN, ixheaac_sub_samp_qmf_window_coeff, 1560 * 4);
  memcpy(hr + XAAC_HROM_SYNCOS, ixheaac_synth_cos_table_kl_4, 16 * 4);
  memcpy(hr + XAAC_HROM_SYNCOS + 16, ixheaac_synth_cos_table_kl_8, 32 * 4);
  memcpy(hr + XAAC_HROM_SYNCOS + 48, ixheaac_synth_cos_table_kl_12, 48 * 4);
  memcpy(hr + XAAC_HROM_SYNCOS + 96, ixheaac_synth_cos_table_kl_16, 64 * 4);
  memcpy(hr + XAAC_HROM_ANACS, ixheaac_analy_cos_sin_table_kl_8, 32 * 4);
  memcpy(hr + XAAC_HROM_ANACS + 32, ixheaac_analy_cos_sin_table_kl_16, 64 * 4);
  memcpy(hr + XAAC_HROM_ANACS + 96, ixheaac_analy_cos_sin_table_kl_24, 96 * 4);
  memcpy(hr + XAAC_HROM_ANACS + 192, ixheaac_analy_cos_sin_table_kl_32, 128 * 4);
  memcpy(hr + XAAC_HROM_COSTRANS, ixheaac_cos_table_trans_qmf, 448 * 4);
  memcpy(hr + XAAC_HROM_FFTTW, ixheaac_twiddle_table_fft_float, 514 * 4);
  memcpy(hr + XAAC_HROM_TW24, ixheaac_twidle_tbl_24, 32 * 4);
  memcpy(hr + XAAC_HROM_TW48, ixheaac_twidle_tbl_48, 64 * 4);
  memcpy(hr + XAAC_HROM_PVCOS, ixheaac_phase_vocoder_cos_table, 64 * 4);
  memcpy(hr + XAAC_HROM_PVSIN, ixheaac_phase_vocoder_sin_table, 64 * 4);
  memcpy(hr + XAAC_HROM_INTERP, ixheaac_hbe_post_anal_proc_interp_coeff, 8 * 4);
  memcpy(hr + XAAC_HROM_SELCASE, ixheaac_sel_case, 40 * 4);
  memcpy(hr + XAAC_HROM_XP2, ixheaac_hbe_x_prod_cos_table_trans_2, 512 * 4);
  memcpy(hr + XAAC_HROM_XP3, ixheaac_hbe_x_prod_cos_table_trans_3, 512 * 4);
  memcpy(hr + XAAC_HROM_XP4, ixheaac_hbe_x_prod_cos_table_trans_4, 512 * 4);
  memcpy(hr + XAAC_HROM_XP41, ixheaac_hbe_x_prod_cos_table_trans_4_1, 512 * 4);
  memcpy(hr + XAAC_HROM_SYN20, ixheaac_synth_cos_table_kl_20, 800 * 4);
  memcpy(hr + XAAC_HROM_ANA40, ixheaac_analy_cos_sin_table_kl_40, 3200 * 4);
  B200(xaac_b200_set_hbe_rom(c, hr, sizeof(hr)), "set_hbe_rom");
  for (int i = 0; i < 6; i++) B200(xaac_b200_dev_alloc(c, 72 * 64 * 4, (void **)&G.e_q[i]), "alloc");
  B200(xaac_b200_dev_alloc(c, 32, (void **)&G.e_bw), "alloc");
  B200(xaac_b200_dev_alloc(c, 640 * 4, (void **)&G.e_ec), "alloc");
  B200(xaac_b200_dev_alloc(c, XAAC_HBE_ST_WORDS * 4, (void **)&G.e_hbe), "alloc");
  B200(xaac_b200_dev_alloc(c, XAAC_EEC_FPAR_WORDS * 4, (void **)&G.e_fpar), "alloc");
  B200(xaac_b200_dev_alloc(c, 4096, (void **)&G.e_tin), "alloc");
  B200(xaac_b200_dev_alloc(c, 8192, (void **)&G.e_out), "alloc");
  B200(xaac_b200_dev_alloc(c, 320 * 4, (void **)&G.e_anal), "alloc");
  B200(xaac_b200_dev_alloc(c, 16, (void **)&G.e_apos), "alloc");
  B200(xaac_b200_dev_alloc(c, 1280 * 4, (void **)&G.e_synth), "alloc");
  B200(xaac_b200_dev_alloc(c, 16, (void **)&G.e_spos), "alloc");
  B200(xaac_b200_dev_alloc(c, 32, (void **)&G.e_patch), "alloc");
  B200(xaac_b200_dev_alloc(c, XAAC_HBE_CFG_WORDS * 4, (void **)&G.e_hbecfg), "alloc");
  B200(xaac_b200_dev_alloc(c, XAAC_EHF_PAR_WORDS * 4, (void **)&G.e_hfpar), "alloc");
  B200(xaac_b200_dev_alloc(c, XAAC_EEC_IPAR_WORDS * 4, (void **)&G.e_ipar), "alloc");
  B200(xaac_b200_dev_alloc(c, 16, (void **)&G.e_rg), "alloc");
  B200(xaac_b200_dev_alloc(c, 32, (void **)&G.e_err), "alloc");
  static float pr[XAAC_FPSROM_WORDS];
  b200_fps_pack_rom(pr, t->ps_tables_ptr, t->ps_tables_ptr->rev_link_delay_ser);
  B200(xaac_b200_set_fps_rom(c, pr, sizeof(pr)), "set_fps_rom");
  B200(xaac_b200_dev_alloc(c, XAAC_FPS_ST_WORDS * 4, (void **)&G.e_ps_state), "alloc");
  B200(xaac_b200_dev_alloc(c, 4096 * 4, (void **)&G.e_ps_left), "alloc");
  B200(xaac_b200_dev_alloc(c, 4096 * 4, (void **)&G.e_ps_right), "alloc");
  B200(xaac_b200_dev_alloc(c, XAAC_FPS_SIDE_WORDS * 4, (void **)&G.e_ps_side), "alloc");
  B200(xaac_b200_dev_alloc(c, 8192, (void **)&G.e_out_r), "alloc");
  B200(xaac_b200_dev_alloc(c, 1280 * 4, (void **)&G.e_synth_r), "alloc");
  B200(xaac_b200_dev_alloc(c, 16, (void **)&G.e_spos_r), "alloc");
  G.have_esbr_rom = 1;
}

static void esbr_pack_hf_par(int32_t *par, const ia_sbr_frame_info_data_struct *fd, const ia_sbr_header_data_struct *hd) {
  const ia_freq_band_data_struct *fb = hd->pstr_freq_band_data;
  memset(par, 0, 4 * XAAC_EHF_PAR_WORDS);
  par[XAAC_EHF_NUM_MF] = fb->num_mf_bands;
  par[XAAC_EHF_NUM_IF] = fb->num_nf_bands;
  par[XAAC_EHF_SB_START] = fb->sub_band_start;
  par[XAAC_EHF_BORDER_FIRST] = fd->str_frame_info_details.border_vec[0];
  par[XAAC_EHF_BORDER_LAST] = fd->str_frame_info_details.border_vec[fd->str_frame_info_details.num_env];
  par[XAAC_EHF_HBE_FLAG] = hd->hbe_flag;
  par[XAAC_EHF_PATCHING_MODE] = fd->sbr_patching_mode;
  par[XAAC_EHF_FS] = hd->out_sampling_freq;
  par[XAAC_EHF_PRE_PROC] = hd->pre_proc_flag;
  par[XAAC_EHF_USF4] = hd->is_usf_4;
  par[XAAC_EHF_MPS_SBR] = fd->mps_sbr_flag;
  par[XAAC_EHF_COV_COUNT] = fd->cov_count;
  for (int i = 0; i < 5; i++) {
    par[XAAC_EHF_INVF + i] = fd->sbr_invf_mode[i];
    par[XAAC_EHF_INVF_PREV + i] = fd->sbr_invf_mode_prev[i];
    par[XAAC_EHF_INVF_TBL + i] = fb->freq_band_tbl_noise[1 + i];
  }
  for (int i = 0; i < 57; i++) par[XAAC_EHF_FMASTER + i] = fb->f_master_tbl[i];
}
static void esbr_pack_ec_ipar(int32_t *ip, const ia_sbr_frame_info_data_struct *fd, const ia_sbr_header_data_struct *hd) {
  const ia_freq_band_data_struct *fb = hd->pstr_freq_band_data;
  const ia_frame_info_struct *fi = &fd->str_frame_info_details;
  memset(ip, 0, 4 * XAAC_EEC_IPAR_WORDS);
  ip[XAAC_EEC_SB_START] = fb->sub_band_start;
  ip[XAAC_EEC_SB_END] = fb->sub_band_end;
  ip[XAAC_EEC_NUM_ENV] = fi->num_env;
  ip[XAAC_EEC_TRANS_ENV] = fi->transient_env;
  ip[XAAC_EEC_SHORT_PREV] = fd->env_short_flag_prev;
  ip[XAAC_EEC_NUM_NOISE_ENV] = fi->num_noise_env;
  ip[XAAC_EEC_NUM_SF_LO] = fb->num_sf_bands[0];
  ip[XAAC_EEC_NUM_SF_HI] = fb->num_sf_bands[1];
  ip[XAAC_EEC_NUM_NF] = fb->num_nf_bands;
  ip[XAAC_EEC_SMOOTHING_MODE] = hd->smoothing_mode;
  ip[XAAC_EEC_INTERPOL_FREQ] = hd->interpol_freq;
  ip[XAAC_EEC_LIMITER_BANDS] = hd->limiter_bands;
  ip[XAAC_EEC_LIMITER_GAINS] = hd->limiter_gains;
  ip[XAAC_EEC_HARM_INDEX] = fd->harm_index;
  ip[XAAC_EEC_PHASE_INDEX] = fd->phase_index;
  ip[XAAC_EEC_START_UP] = hd->esbr_start_up;
  ip[XAAC_EEC_RESET] = fd->reset_flag;
  ip[XAAC_EEC_SBR_MODE] = fd->sbr_mode;
  ip[XAAC_EEC_USF4] = hd->is_usf_4;
  ip[XAAC_EEC_PATCHING_CHANGED] = fd->sbr_patching_mode != fd->prev_sbr_patching_mode;
  for (int i = 0; i < 9; i++) ip[XAAC_EEC_BORDER + i] = fi->border_vec[i];
  for (int i = 0; i < 8; i++) ip[XAAC_EEC_FREQ_RES + i] = fi->freq_res[i];
  for (int i = 0; i < 3; i++) ip[XAAC_EEC_NOISE_BORDER + i] = fi->noise_border_vec[i];
  for (int i = 0; i < 8; i++) ip[XAAC_EEC_INTER_TES + i] = fd->inter_temp_shape_mode[i];
  for (int i = 0; i < 4; i++) ip[XAAC_EEC_GATE_MODE + i] = fd->gate_mode[i];
  for (int i = 0; i < 52; i++) ip[XAAC_EEC_LIM_TABLE + i] = fd->lim_table[i / 13][i % 13];
  for (int i = 0; i < 6; i++) ip[XAAC_EEC_TBL_NOISE + i] = fb->freq_band_tbl_noise[i];
  for (int i = 0; i < 29; i++) ip[XAAC_EEC_TBL_LO + i] = fb->freq_band_tbl_lo[i];
  for (int i = 0; i < 57; i++) ip[XAAC_EEC_TBL_HI + i] = fb->freq_band_tbl_hi[i];
  for (int i = 0; i < 56; i++) ip[XAAC_EEC_ADD_HARM + i] = fd->add_harmonics[i];
  memcpy(ip + XAAC_EEC_HARM_PREV, fd->harm_flag_prev, 64);
}

static WORD32 esbr_dec_b200(xaac_b200_ctx *c, ia_sbr_dec_struct *d, ia_sbr_header_data_struct *hd,
                            ia_sbr_frame_info_data_struct *fd, ia_sbr_tables_struct *t, ia_pvc_data_struct *pvc,
                            ia_ps_dec_struct *ps, VOID *self, int *done) {
  *done = 0;
  esbr_rom_once(c, t);
  /* mono + PS element (sbr_dec.c:976-1001): the right channel leaves through the second channel's synthesis bank */
  ia_sbr_qmf_filter_bank_struct *yr = NULL;
  static float ps_side[XAAC_FPS_SIDE_WORDS], ps_st[XAAC_FPS_ST_WORDS];
  static b200_fps_commit_rec ps_cm;
  int32_t spos_r[2] = {0, 0};
  if (ps) {
    if (!self) return 0;
    yr = &((ia_handle_sbr_dec_inst_struct)self)->pstr_sbr_channel[1]->str_sbr_dec.str_synthesis_qmf_bank;
    if (yr->no_channels != 64) return 0;
    if (b200_fps_side(ps_side, &ps_cm, ps, t->ps_tables_ptr, hd->pstr_freq_band_data->sub_band_end) != 0) return 0;
    b200_fps_pack_state(ps_st, ps);
    spos_r[0] = yr->ixheaacd_drc_offset;
    spos_r[1] = (int32_t)(yr->filter_pos_syn_32 - yr->p_filter_32);
    if (spos_r[1] < 0 || spos_r[1] > 640) return 0;
  }
  const int hbe = hd->hbe_flag != 0;
  ia_esbr_hbe_txposer_struct *tx = d->p_hbe_txposer;
  ia_sbr_qmf_filter_bank_struct *a = &d->str_codec_qmf_bank, *y = &d->str_synthesis_qmf_bank;
  WORD32 *qc = (WORD32 *)t->qmf_dec_tables_ptr->esbr_qmf_c;
  const int rows = hbe ? 72 : 40;
  static int32_t hf_par[XAAC_EHF_PAR_WORDS], ipar[XAAC_EEC_IPAR_WORDS], hcfg[XAAC_HBE_CFG_WORDS], rg[4], apos[2], spos[2], patch[8];
  static float fpar[XAAC_EEC_FPAR_WORDS], ec[640], hst[XAAC_HBE_ST_WORDS], bw[6];
  int32_t err[6];
  if (hbe) {
    if (!tx) return 0;
    if (tx->ixheaacd_cmplx_anal_fft == NULL) { /* hbe_trans.c:240-248: the reference re-initialises inside the call */
      WORD32 e = ixheaacd_qmf_hbe_data_reinit(tx, hd->pstr_freq_band_data->freq_band_table, hd->pstr_freq_band_data->num_sf_bands,
                                              hd->is_usf_4);
      if (e) return 0; /* let the reference report it */
    }
    const int S = tx->synth_size;
    if (S < 1 || S > 20 || tx->no_bins != 32) return 0;
    memset(hcfg, 0, sizeof(hcfg));
    hcfg[XAAC_HBE_SYNTH_SIZE] = S;
    hcfg[XAAC_HBE_K_START] = tx->k_start;
    hcfg[XAAC_HBE_START_BAND] = tx->start_band;
    hcfg[XAAC_HBE_END_BAND] = tx->end_band;
    hcfg[XAAC_HBE_MAX_STRETCH] = tx->max_stretch;
    hcfg[XAAC_HBE_PITCH] = fd->pitch_in_bins;
    hcfg[XAAC_HBE_USF4] = tx->upsamp_4_flag;
    for (int i = 0; i < 6; i++) hcfg[XAAC_HBE_XOVER + i] = tx->x_over_qmf[i];
    memset(hst, 0, sizeof(hst));
    memcpy(hst + XAAC_HBE_ST_TAIL, tx->ptr_input_buf + tx->no_bins * S, S * 4);
    memcpy(hst + XAAC_HBE_ST_SYNTH, tx->synth_buf, 18 * S * 4);
    memcpy(hst + XAAC_HBE_ST_ANAL, tx->analy_buf, 18 * S * 4);
    for (int r = 0; r < 12; r++) memcpy(hst + XAAC_HBE_ST_QIN + 128 * r, tx->qmf_in_buf[16 + r], 512);
    for (int r = 0; r < 10; r++) memcpy(hst + XAAC_HBE_ST_QOUT + 128 * r, tx->qmf_out_buf[32 + r], 512);
  }
  if (hbe && !hd->usac_flag) {
    /* sbr_dec.c:868-874 (legacy streams): after the history shift the reference clears (64 - qmf_sb_prev) BYTES — not floats —
     * from band qmf_sb_prev of rows 2..7; those are rows 34..39 before the shift, which the kernel performs */
    const int q = hd->pstr_freq_band_data->qmf_sb_prev;
    if (q >= 0 && q < 64)
      for (int i = 2; i < 8; i++) {
        memset(&d->qmf_buf_real[32 + i][q], 0, (size_t)(64 - q));
        memset(&d->qmf_buf_imag[32 + i][q], 0, (size_t)(64 - q));
      }
  }
  esbr_pack_hf_par(hf_par, fd, hd);
  esbr_pack_ec_ipar(ipar, fd, hd);
  memset(fpar, 0, sizeof(fpar));
  memcpy(fpar + XAAC_EEC_SFB_NRG, fd->flt_env_sf_arr, 448 * 4);
  memcpy(fpar + XAAC_EEC_NOISE_FLOOR, fd->flt_noise_floor, 10 * 4);
  memcpy(ec, fd->e_gain, 320 * 4);
  memcpy(ec + 320, fd->noise_buf, 320 * 4);
  memcpy(bw, fd->bw_array_prev, sizeof(bw));
  patch[0] = fd->patch_param.num_patches;
  for (int i = 0; i < 7; i++) patch[1 + i] = fd->patch_param.start_subband[i];
  rg[0] = hd->pstr_freq_band_data->qmf_sb_prev;
  rg[1] = hd->pstr_freq_band_data->sub_band_start;
  rg[2] = 2 * fd->str_frame_info_details.border_vec[0];
  rg[3] = 0;
  apos[0] = (int32_t)(a->state_new_samples_pos_low_32 - a->anal_filter_states_32);
  apos[1] = (int32_t)(a->filter_pos_32 - qc);
  spos[0] = y->ixheaacd_drc_offset;
  spos[1] = (int32_t)(y->filter_pos_syn_32 - y->p_filter_32);
  if (apos[0] < 0 || apos[0] >= 320 || apos[1] < 0 || apos[1] > 640 || spos[1] < 0 || spos[1] > 640) return 0;
  float *src[6] = {&d->qmf_buf_real[0][0], &d->qmf_buf_imag[0][0], &d->sbr_qmf_out_real[0][0], &d->sbr_qmf_out_imag[0][0],
                   &d->ph_vocod_qmf_real[0][0], &d->ph_vocod_qmf_imag[0][0]};
  for (int i = 0; i < (hbe ? 6 : 4); i++)
    B200(xaac_b200_h2d(c, G.e_q[i], src[i], (size_t)(i < 2 ? rows : 40) * 64 * 4), "h2d qmf");
  B200(xaac_b200_h2d(c, G.e_anal, a->anal_filter_states_32, 320 * 4), "h2d");
  B200(xaac_b200_h2d(c, G.e_apos, apos, 8), "h2d");
  B200(xaac_b200_h2d(c, G.e_synth, y->filter_states_32, 1280 * 4), "h2d");
  B200(xaac_b200_h2d(c, G.e_spos, spos, 8), "h2d");
  B200(xaac_b200_h2d(c, G.e_bw, bw, 24), "h2d");
  B200(xaac_b200_h2d(c, G.e_patch, patch, 32), "h2d");
  B200(xaac_b200_h2d(c, G.e_ec, ec, sizeof(ec)), "h2d");
  B200(xaac_b200_h2d(c, G.e_hfpar, hf_par, sizeof(hf_par)), "h2d");
  B200(xaac_b200_h2d(c, G.e_ipar, ipar, sizeof(ipar)), "h2d");
  B200(xaac_b200_h2d(c, G.e_fpar, fpar, sizeof(fpar)), "h2d");
  B200(xaac_b200_h2d(c, G.e_rg, rg, 16), "h2d");
  B200(xaac_b200_h2d(c, G.e_tin, d->time_sample_buf, 4096), "h2d");
  xaac_b200_esbr_hbe_state_view v;
  memset(&v, 0, sizeof(v));
  v.base.qmf_re = G.e_q[0]; v.base.qmf_im = G.e_q[1]; v.base.out_re = G.e_q[2]; v.base.out_im = G.e_q[3];
  v.base.anal_states = G.e_anal; v.base.anal_pos = G.e_apos; v.base.synth_states = G.e_synth; v.base.synth_pos = G.e_spos;
  v.base.bw_prev = G.e_bw; v.base.patch = G.e_patch; v.base.ec_state = G.e_ec;
  v.pv_re = G.e_q[4]; v.pv_im = G.e_q[5]; v.hbe_state = G.e_hbe;
  memset(err, 0, sizeof(err));
  B200(xaac_b200_h2d(c, G.e_err, err, sizeof(err)), "h2d");
  if (hbe) {
    B200(xaac_b200_h2d(c, G.e_hbe, hst, sizeof(hst)), "h2d");
    B200(xaac_b200_h2d(c, G.e_hbecfg, hcfg, sizeof(hcfg)), "h2d");
  }
  /* the stage in its two halves (include/xaac_b200.h): between them the host does what only the host can do on reset frames and on
   * frames where sbr_patching_mode changes — rebuild the limiter tables (ixheaacd_createlimiterbands, esbr_envcal.c:169-190) from the
   * patch table the HF generator of this very frame has produced */
  xaac_b200_esbr_ps_view pv;
  memset(&pv, 0, sizeof(pv));
  if (ps) {
    pv.ps_state = G.e_ps_state; pv.left = G.e_ps_left; pv.right = G.e_ps_right; pv.synth_states_r = G.e_synth_r;
    pv.synth_pos_r = G.e_spos_r;
    B200(xaac_b200_h2d(c, G.e_ps_state, ps_st, sizeof(ps_st)), "h2d ps");
    B200(xaac_b200_h2d(c, G.e_ps_side, ps_side, sizeof(ps_side)), "h2d ps");
    B200(xaac_b200_h2d(c, G.e_synth_r, yr->filter_states_32, 1280 * 4), "h2d ps");
    B200(xaac_b200_h2d(c, G.e_spos_r, spos_r, 8), "h2d ps");
  }
  if (!hbe) { v.pv_re = NULL; v.pv_im = NULL; v.hbe_state = NULL; }
  B200(xaac_b200_esbr_dec_front_dev(c, &v, G.e_tin, NULL, hbe ? G.e_hbecfg : NULL, G.e_hfpar, G.e_err, 1, NULL), "esbr_dec_front_dev");
  if (fd->reset_flag || fd->sbr_patching_mode != fd->prev_sbr_patching_mode) {
    B200(xaac_b200_d2h(c, err, G.e_err, sizeof(err)), "d2h err");
    memcpy(G.last_err, err, sizeof(err));
    for (int i = 0; i < 6; i++)
      if (err[i] == -2) return 0;
    for (int i = 0; i < 6; i++)
      if (err[i] != 0) { *done = 1; return err[i]; }
    B200(xaac_b200_d2h(c, patch, G.e_patch, 32), "d2h patch");
    fd->patch_param.num_patches = patch[0]; /* what ixheaacd_generate_hf leaves there (the reference recomputes the same on a fallback) */
    for (int i = 0; i < 7; i++) fd->patch_param.start_subband[i] = patch[1 + i];
    if (ixheaacd_createlimiterbands(fd->lim_table, fd->gate_mode, hd->pstr_freq_band_data->freq_band_tbl_lo,
                                    hd->pstr_freq_band_data->num_sf_bands[LOW], hbe ? tx->x_over_qmf : NULL, fd->sbr_patching_mode,
                                    hd->is_usf_4, &fd->patch_param, 0)) {
      *done = 1;
      return IA_FATAL_ERROR;
    }
    for (int i = 0; i < 4; i++) ipar[XAAC_EEC_GATE_MODE + i] = fd->gate_mode[i];
    for (int i = 0; i < 52; i++) ipar[XAAC_EEC_LIM_TABLE + i] = fd->lim_table[i / 13][i % 13];
    ipar[XAAC_EEC_LIM_REBUILT] = 1;
    B200(xaac_b200_h2d(c, G.e_ipar, ipar, sizeof(ipar)), "h2d");
    G.n_esbr_rebuilt++;
  }
  B200(xaac_b200_esbr_dec_back_dev(c, &v, ps ? &pv : NULL, G.e_ipar, G.e_fpar, G.e_rg, ps ? G.e_ps_side : NULL, G.e_out,
                                   ps ? G.e_out_r : NULL, NULL, 1, G.e_err, 1, NULL), "esbr_dec_back_dev");
  B200(xaac_b200_d2h(c, err, G.e_err, sizeof(err)), "d2h err");
  memcpy(G.last_err, err, sizeof(err));
  for (int i = 0; i < 6; i++)
    if (err[i] == -2) return 0; /* outside the kernels' subset: nothing on the host has been touched yet */
  *done = 1;
  for (int i = 0; i < 6; i++)
    if (err[i] != 0) return err[i]; /* the reference's own failure codes */
  /* ---- results and state back into the reference's structs ---- */
  for (int i = 0; i < (hbe ? 6 : 4); i++)
    B200(xaac_b200_d2h(c, src[i], G.e_q[i], (size_t)(i < 2 ? rows : 40) * 64 * 4), "d2h qmf");
  B200(xaac_b200_d2h(c, a->anal_filter_states_32, G.e_anal, 320 * 4), "d2h");
  B200(xaac_b200_d2h(c, apos, G.e_apos, 8), "d2h");
  B200(xaac_b200_d2h(c, y->filter_states_32, G.e_synth, 1280 * 4), "d2h");
  B200(xaac_b200_d2h(c, spos, G.e_spos, 8), "d2h");
  B200(xaac_b200_d2h(c, bw, G.e_bw, 24), "d2h");
  B200(xaac_b200_d2h(c, patch, G.e_patch, 32), "d2h");
  B200(xaac_b200_d2h(c, ec, G.e_ec, sizeof(ec)), "d2h");
  static int32_t ipar_in[XAAC_EEC_IPAR_WORDS];
  memcpy(ipar_in, ipar, sizeof(ipar));
  B200(xaac_b200_d2h(c, ipar, G.e_ipar, sizeof(ipar)), "d2h");
  if (ps) {
    B200(xaac_b200_d2h(c, ps->time_sample_buf[0], G.e_out, 8192), "d2h out");
    B200(xaac_b200_d2h(c, ps->time_sample_buf[1], G.e_out_r, 8192), "d2h out");
    B200(xaac_b200_d2h(c, ps_st, G.e_ps_state, sizeof(ps_st)), "d2h ps");
    b200_fps_unpack_state(ps_st, ps);
    b200_fps_commit(ps, &ps_cm, t->ps_tables_ptr);
    B200(xaac_b200_d2h(c, yr->filter_states_32, G.e_synth_r, 1280 * 4), "d2h ps");
    B200(xaac_b200_d2h(c, spos_r, G.e_spos_r, 8), "d2h ps");
    yr->esbr_cos_twiddle = (WORD32 *)t->qmf_dec_tables_ptr->esbr_sin_cos_twiddle_l64; /* sbr_dec.c:567-580, second call */
    yr->esbr_alt_sin_twiddle = (WORD32 *)t->qmf_dec_tables_ptr->esbr_alt_sin_twiddle_l64;
    yr->p_filter_32 = qc;
    yr->filter_pos_syn_32 = qc + spos_r[1];
    yr->ixheaacd_drc_offset = spos_r[0];
    ((ia_sbr_frame_info_data_struct *)((ia_handle_sbr_dec_inst_struct)self)->frame_buffer[1])->reset_flag = 0; /* sbr_dec.c:659 */
  } else {
    B200(xaac_b200_d2h(c, d->time_sample_buf, G.e_out, 8192), "d2h out");
  }
  a->usb = a->no_channels; /* sbr_dec.c:238 */
  a->state_new_samples_pos_low_32 = a->anal_filter_states_32 + apos[0];
  a->filter_pos_32 = qc + apos[1];
  y->esbr_cos_twiddle = (WORD32 *)t->qmf_dec_tables_ptr->esbr_sin_cos_twiddle_l64; /* sbr_dec.c:567-580 */
  y->esbr_alt_sin_twiddle = (WORD32 *)t->qmf_dec_tables_ptr->esbr_alt_sin_twiddle_l64;
  y->p_filter_32 = qc;
  y->filter_pos_syn_32 = qc + spos[1];
  y->ixheaacd_drc_offset = spos[0];
  memcpy(fd->bw_array_prev, bw, sizeof(bw));
  fd->patch_param.num_patches = patch[0];
  for (int i = 0; i < 7; i++) fd->patch_param.start_subband[i] = patch[1 + i];
  memcpy(fd->e_gain, ec, 320 * 4);
  memcpy(fd->noise_buf, ec + 320, 320 * 4);
  /* ixheaacd_sbr_env_calc's epilogue (esbr_envcal.c:861-908) */
  const int sbs = hd->pstr_freq_band_data->sub_band_start;
  memcpy(fd->harm_flag_varlen_prev, fd->harm_flag_prev, 64);
  memcpy(fd->harm_flag_prev, ipar + XAAC_EEC_HARM_PREV, 64);
  for (int i = 0; i < 64; i++) fd->harm_flag_varlen[i] = i >= sbs ? fd->harm_flag_prev[i] : 0;
  fd->env_short_flag_prev = ipar[XAAC_EEC_SHORT_PREV];
  memcpy(&fd->str_frame_info_prev, &fd->str_frame_info_details, sizeof(ia_frame_info_struct));
  if (fd->str_frame_info_details.num_env == 1) fd->var_len_id_prev = 0;
  else if (fd->str_frame_info_details.num_env == 2) fd->var_len_id_prev = 1;
  {
    const int nnf = hd->pstr_freq_band_data->num_nf_bands;
    for (int i = 0; i < nnf; i++)
      fd->prev_noise_level[i] = fd->flt_noise_floor[(fd->str_frame_info_details.num_noise_env - 1) * nnf + i];
  }
  fd->harm_index = ipar[XAAC_EEC_HARM_INDEX];
  fd->phase_index = ipar[XAAC_EEC_PHASE_INDEX];
  hd->esbr_start_up = ipar[XAAC_EEC_START_UP];
  /* the rest of the stage's bookkeeping (sbr_dec.c:931-1003) */
  pvc->pvc_rate = hd->upsamp_fac;
  pvc->prev_pvc_flg = 0;
  pvc->prev_first_bnd_idx = hd->pstr_freq_band_data->sub_band_start;
  pvc->prev_pvc_rate = pvc->pvc_rate;
  fd->pstr_sbr_header = hd;
  d->band_count = hd->pstr_freq_band_data->sub_band_end;
  fd->reset_flag = 0;
  fd->prev_sbr_mode = fd->sbr_mode;
  fd->prev_sbr_patching_mode = fd->sbr_patching_mode; /* esbr_envcal.c:189 */
  if (hbe) {
    const int S = tx->synth_size;
    B200(xaac_b200_d2h(c, hst, G.e_hbe, sizeof(hst)), "d2h hbe");
    memcpy(tx->ptr_input_buf + tx->no_bins * S, hst + XAAC_HBE_ST_TAIL, S * 4);
    memcpy(tx->synth_buf, hst + XAAC_HBE_ST_SYNTH, 18 * S * 4);
    memcpy(tx->analy_buf, hst + XAAC_HBE_ST_ANAL, 18 * S * 4);
    for (int r = 0; r < 12; r++) memcpy(tx->qmf_in_buf[16 + r], hst + XAAC_HBE_ST_QIN + 128 * r, 512);
    for (int r = 0; r < 10; r++) memcpy(tx->qmf_out_buf[32 + r], hst + XAAC_HBE_ST_QOUT + 128 * r, 512);
    for (int r = 42; r < 64; r++) memset(tx->qmf_out_buf[r], 0, 512);
  }
  (void)ipar_in;
  return 0;
}

/* apply_processing = 0: the frames before the first SBR header of a stream only pass through the banks (sbr_dec.c:836-878, 964-1003) */
static WORD32 esbr_bypass_b200(xaac_b200_ctx *c, ia_sbr_dec_struct *d, ia_sbr_header_data_struct *hd,
                               ia_sbr_frame_info_data_struct *fd, ia_sbr_tables_struct *t, ia_ps_dec_struct *ps, VOID *self,
                               int *done) {
  *done = 0;
  esbr_rom_once(c, t);
  const int hbe = hd->hbe_flag != 0;
  const int rows = hbe ? 72 : 40;
  ia_sbr_qmf_filter_bank_struct *a = &d->str_codec_qmf_bank, *y = &d->str_synthesis_qmf_bank, *yr = NULL;
  WORD32 *qc = (WORD32 *)t->qmf_dec_tables_ptr->esbr_qmf_c;
  int32_t rg[4], apos[2], spos[2], spos_r[2] = {0, 0}, err[6];
  if (ps) {
    if (!self) return 0;
    yr = &((ia_handle_sbr_dec_inst_struct)self)->pstr_sbr_channel[1]->str_sbr_dec.str_synthesis_qmf_bank;
    if (yr->no_channels != 64) return 0;
    spos_r[0] = yr->ixheaacd_drc_offset;
    spos_r[1] = (int32_t)(yr->filter_pos_syn_32 - yr->p_filter_32);
    if (spos_r[1] < 0 || spos_r[1] > 640) return 0;
  }
  apos[0] = (int32_t)(a->state_new_samples_pos_low_32 - a->anal_filter_states_32);
  apos[1] = (int32_t)(a->filter_pos_32 - qc);
  spos[0] = y->ixheaacd_drc_offset;
  spos[1] = (int32_t)(y->filter_pos_syn_32 - y->p_filter_32);
  if (apos[0] < 0 || apos[0] >= 320 || apos[1] < 0 || apos[1] > 640 || spos[1] < 0 || spos[1] > 640) return 0;
  if (hbe && !hd->usac_flag) { /* sbr_dec.c:868-874, see esbr_dec_b200 */
    const int q = hd->pstr_freq_band_data->qmf_sb_prev;
    if (q >= 0 && q < 64)
      for (int i = 2; i < 8; i++) {
        memset(&d->qmf_buf_real[32 + i][q], 0, (size_t)(64 - q));
        memset(&d->qmf_buf_imag[32 + i][q], 0, (size_t)(64 - q));
      }
  }
  rg[0] = 0; rg[1] = hd->pstr_freq_band_data->sub_band_start; rg[2] = 0; rg[3] = 0; /* sbr_dec.c:305-318, 380: stop_border 0 */
  float *src[6] = {&d->qmf_buf_real[0][0], &d->qmf_buf_imag[0][0], &d->sbr_qmf_out_real[0][0], &d->sbr_qmf_out_imag[0][0],
                   &d->ph_vocod_qmf_real[0][0], &d->ph_vocod_qmf_imag[0][0]};
  for (int i = 0; i < (hbe ? 6 : 2); i++)
    if (i < 2 || i > 3) B200(xaac_b200_h2d(c, G.e_q[i], src[i], (size_t)(i < 2 ? rows : 40) * 64 * 4), "h2d qmf");
  B200(xaac_b200_h2d(c, G.e_anal, a->anal_filter_states_32, 320 * 4), "h2d");
  B200(xaac_b200_h2d(c, G.e_apos, apos, 8), "h2d");
  B200(xaac_b200_h2d(c, G.e_synth, y->filter_states_32, 1280 * 4), "h2d");
  B200(xaac_b200_h2d(c, G.e_spos, spos, 8), "h2d");
  B200(xaac_b200_h2d(c, G.e_rg, rg, 16), "h2d");
  B200(xaac_b200_h2d(c, G.e_tin, d->time_sample_buf, 4096), "h2d");
  memset(err, 0, sizeof(err));
  B200(xaac_b200_h2d(c, G.e_err, err, sizeof(err)), "h2d");
  xaac_b200_esbr_hbe_state_view v;
  xaac_b200_esbr_ps_view pv;
  memset(&v, 0, sizeof(v));
  memset(&pv, 0, sizeof(pv));
  v.base.qmf_re = G.e_q[0]; v.base.qmf_im = G.e_q[1]; v.base.out_re = G.e_q[2]; v.base.out_im = G.e_q[3];
  v.base.anal_states = G.e_anal; v.base.anal_pos = G.e_apos; v.base.synth_states = G.e_synth; v.base.synth_pos = G.e_spos;
  if (hbe) { v.pv_re = G.e_q[4]; v.pv_im = G.e_q[5]; v.hbe_state = G.e_hbe; }
  if (ps) {
    pv.synth_states_r = G.e_synth_r; pv.synth_pos_r = G.e_spos_r;
    B200(xaac_b200_h2d(c, G.e_synth_r, yr->filter_states_32, 1280 * 4), "h2d ps");
    B200(xaac_b200_h2d(c, G.e_spos_r, spos_r, 8), "h2d ps");
  }
  B200(xaac_b200_esbr_dec_bypass_dev(c, &v, ps ? &pv : NULL, G.e_tin, NULL, G.e_rg, G.e_out, ps ? G.e_out_r : NULL, NULL, 1, G.e_err, 1,
                                     NULL), "esbr_dec_bypass_dev");
  B200(xaac_b200_d2h(c, err, G.e_err, sizeof(err)), "d2h err");
  memcpy(G.last_err, err, sizeof(err));
  for (int i = 0; i < 6; i++)
    if (err[i] == -2) return 0;
  *done = 1;
  for (int i = 0; i < 6; i++)
    if (err[i] != 0) return err[i];
  for (int i = 0; i < (hbe ? 6 : 2); i++)
    if (i < 2 || i > 3) B200(xaac_b200_d2h(c, src[i], G.e_q[i], (size_t)(i < 2 ? rows : 40) * 64 * 4), "d2h qmf");
  for (int i = 0; i < 64; i++) { /* sbr_dec.c:964-969 */
    memset(d->sbr_qmf_out_real[i], 0, 64 * sizeof(FLOAT32));
    memset(d->sbr_qmf_out_imag[i], 0, 64 * sizeof(FLOAT32));
  }
  B200(xaac_b200_d2h(c, a->anal_filter_states_32, G.e_anal, 320 * 4), "d2h");
  B200(xaac_b200_d2h(c, apos, G.e_apos, 8), "d2h");
  B200(xaac_b200_d2h(c, y->filter_states_32, G.e_synth, 1280 * 4), "d2h");
  B200(xaac_b200_d2h(c, spos, G.e_spos, 8), "d2h");
  a->usb = a->no_channels;
  a->state_new_samples_pos_low_32 = a->anal_filter_states_32 + apos[0];
  a->filter_pos_32 = qc + apos[1];
  y->esbr_cos_twiddle = (WORD32 *)t->qmf_dec_tables_ptr->esbr_sin_cos_twiddle_l64;
  y->esbr_alt_sin_twiddle = (WORD32 *)t->qmf_dec_tables_ptr->esbr_alt_sin_twiddle_l64;
  y->p_filter_32 = qc;
  y->filter_pos_syn_32 = qc + spos[1];
  y->ixheaacd_drc_offset = spos[0];
  if (ps) {
    B200(xaac_b200_d2h(c, ps->time_sample_buf[0], G.e_out, 8192), "d2h out");
    B200(xaac_b200_d2h(c, ps->time_sample_buf[1], G.e_out_r, 8192), "d2h out");
    B200(xaac_b200_d2h(c, yr->filter_states_32, G.e_synth_r, 1280 * 4), "d2h ps");
    B200(xaac_b200_d2h(c, spos_r, G.e_spos_r, 8), "d2h ps");
    yr->esbr_cos_twiddle = (WORD32 *)t->qmf_dec_tables_ptr->esbr_sin_cos_twiddle_l64;
    yr->esbr_alt_sin_twiddle = (WORD32 *)t->qmf_dec_tables_ptr->esbr_alt_sin_twiddle_l64;
    yr->p_filter_32 = qc;
    yr->filter_pos_syn_32 = qc + spos_r[1];
    yr->ixheaacd_drc_offset = spos_r[0];
    ((ia_sbr_frame_info_data_struct *)((ia_handle_sbr_dec_inst_struct)self)->frame_buffer[1])->reset_flag = 0;
  } else {
    B200(xaac_b200_d2h(c, d->time_sample_buf, G.e_out, 8192), "d2h out");
  }
  d->band_count = hd->pstr_freq_band_data->sub_band_end;
  fd->reset_flag = 0;
  fd->prev_sbr_mode = fd->sbr_mode;
  return 0;
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
  if (c && ptr_header_data->enh_sbr && audio_object_type != AOT_ER_AAC_ELD && audio_object_type != AOT_ER_AAC_LD) {
    /* float eSBR branch: a USAC channel, or a legacy HE-AAC channel decoded in the reference's default (eSBR) mode, where the
     * harmonic transposer is forced on (decoder/ixheaacd_sbrdecoder.c:400-403) */
    int tes = 0;
    for (int i = 0; i < 8; i++) tes |= ptr_frame_data->inter_temp_shape_mode[i];
    const int with_ps = ptr_header_data->channel_mode == PS_STEREO || ptr_header_data->enh_sbr_ps;
    const int ok = (!low_pow_flag || !ptr_header_data->usac_flag) && !ldmps_present && !drc_on &&
                   !heaac_mps_present && !ec_flag && (ptr_header_data->usac_flag || ptr_header_data->hbe_flag) &&
                   ptr_header_data->num_time_slots == 16 && ptr_sbr_dec->str_codec_qmf_bank.no_channels == 32 &&
                   ptr_sbr_dec->str_synthesis_qmf_bank.no_channels == 64 && (!with_ps || (ptr_ps_dec != NULL && self != NULL)) &&
                   ptr_frame_data->stereo_config_idx <= 0 && !ptr_frame_data->mps_sbr_flag && !ptr_header_data->is_usf_4 &&
                   (!apply_processing ||
                    (ptr_frame_data->sbr_mode == ORIG_SBR &&
                     (ptr_frame_data->str_frame_info_details.num_noise_env == 1 ||
                      ptr_frame_data->str_frame_info_details.num_noise_env == 2)));
    if (ok) {
      int done = 0;
      WORD32 r = apply_processing
                     ? esbr_dec_b200(c, ptr_sbr_dec, ptr_header_data, ptr_frame_data, sbr_tables_ptr, ptr_pvc_data_str,
                                     with_ps ? ptr_ps_dec : NULL, self, &done)
                     : esbr_bypass_b200(c, ptr_sbr_dec, ptr_header_data, ptr_frame_data, sbr_tables_ptr, with_ps ? ptr_ps_dec : NULL,
                                        self, &done);
      if (done) {
        if (apply_processing && tes) G.n_esbr_tes++;
        if (!apply_processing) G.n_esbr_bypass++; else if (with_ps) G.n_esbr_ps++; else if (ptr_header_data->hbe_flag) G.n_esbr_hbe++; else G.n_esbr++;
        return r;
      }
    }
    G.n_esbr_ref++;
    if (G.stats > 1)
      fprintf(stderr, "[ixheaacd_b200] eSBR frame left to the reference: eligible %d (apply %d low_pow %d usac %d hbe %d slots %d ps %d "
              "sbr_mode %d usf4 %d pre_proc %d tes %d noise_env %d reset %d) err %d %d %d %d %d %d\n", ok, (int)apply_processing,
              (int)low_pow_flag, (int)ptr_header_data->usac_flag, (int)ptr_header_data->hbe_flag, (int)ptr_header_data->num_time_slots,
              with_ps, (int)ptr_frame_data->sbr_mode, (int)ptr_header_data->is_usf_4, (int)ptr_header_data->pre_proc_flag, tes,
              (int)ptr_frame_data->str_frame_info_details.num_noise_env, (int)ptr_frame_data->reset_flag, G.last_err[0], G.last_err[1],
              G.last_err[2], G.last_err[3], G.last_err[4], G.last_err[5]);
    return __real_ixheaacd_sbr_dec(ptr_sbr_dec, ptr_time_data, ptr_header_data, ptr_frame_data, ptr_frame_data_prev, ptr_ps_dec,
                                   ptr_qmf_synth_bank_r, ptr_sbr_sf_r, apply_processing, low_pow_flag, ptr_work_buf_core,
                                   sbr_tables_ptr, pstr_common_tables, ch_fac, ptr_pvc_data_str, drc_on, drc_sbr_factors,
                                   audio_object_type, ldmps_present, self, heaac_mps_present, ec_flag);
  }
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

/* ================================ ixheaacd_channel_pair_process ================================ */
IA_ERRORCODE __wrap_ixheaacd_channel_pair_process(ia_aac_dec_channel_info_struct *ptr_aac_dec_channel_info[CHANNELS], WORD32 num_ch,
                                                  ia_aac_dec_tables_struct *ptr_aac_tables, WORD32 total_channels,
                                                  WORD32 object_type, WORD32 aac_spect_data_resil_flag,
                                                  WORD32 aac_sf_data_resil_flag, WORD32 *in_data, WORD32 *out_data, void *self_ptr) {
  xaac_b200_ctx *c = b200_ctx();
  static uint8_t rec[XAAC_SPS_BYTES];
  const int eligible = c && (object_type == AOT_AAC_LC || object_type == AOT_SBR || object_type == AOT_PS) && total_channels <= 2 &&
                       num_ch >= 1 && num_ch <= 2 && !aac_spect_data_resil_flag && !aac_sf_data_resil_flag;
  if (eligible && b200_sps_pack(rec, ptr_aac_dec_channel_info, num_ch, ptr_aac_tables) == 0) {
    int32_t err = 0;
    if (!G.have_block_rom) {
      B200(xaac_b200_set_block_rom(c, ptr_aac_tables->pstr_block_tables, 620), "set_block_rom");
      B200(xaac_b200_dev_alloc(c, XAAC_SPS_BYTES, (void **)&G.d_sps), "alloc");
      B200(xaac_b200_dev_alloc(c, 8192, (void **)&G.d_sps_spec), "alloc");
      B200(xaac_b200_dev_alloc(c, 16, (void **)&G.d_sps_seed), "alloc");
      G.have_block_rom = 1;
    }
    B200(xaac_b200_h2d(c, G.d_sps, rec, sizeof(rec)), "h2d sps");
    for (int ch = 0; ch < num_ch; ch++)
      B200(xaac_b200_h2d(c, G.d_sps_spec + 1024 * ch, ptr_aac_dec_channel_info[ch]->ptr_spec_coeff, 4096), "h2d spec");
    int any_pns = 0;
    for (int ch = 0; ch < num_ch; ch++) any_pns |= ptr_aac_dec_channel_info[ch]->str_pns_info.pns_active != 0;
    ia_pns_rand_vec_struct *rnd = ptr_aac_dec_channel_info[0]->pstr_pns_rand_vec_data;
    int32_t seed = rnd->current_seed;
    if (any_pns) B200(xaac_b200_h2d(c, G.d_sps_seed, &seed, 4), "h2d seed");
    B200(xaac_b200_aac_spectral_dev(c, G.d_sps_spec, G.d_sps, any_pns ? G.d_sps_seed : NULL, G.d_err, 1, NULL), "aac_spectral_dev");
    B200(xaac_b200_d2h(c, &err, G.d_err, 4), "d2h err");
    if (err == 0) {
      for (int ch = 0; ch < num_ch; ch++)
        B200(xaac_b200_d2h(c, ptr_aac_dec_channel_info[ch]->ptr_spec_coeff, G.d_sps_spec + 1024 * ch, 4096), "d2h spec");
      /* ixheaacd_pns_process still counts the frames of the first channel (pns_js_thumb.c:196-198) */
      rnd->pns_frame_number++;
      if (any_pns) {
        B200(xaac_b200_d2h(c, &seed, G.d_sps_seed, 4), "d2h seed");
        rnd->current_seed = seed;
        /* what else the reference leaves in the structs: the mask / correlation flags as ixheaacd_map_ms_mask_pns rewrites them
         * (its own code, run after the fact on the untouched side info); random_vector is scratch of the call */
        if (num_ch > 1 && ptr_aac_dec_channel_info[0]->common_window) ixheaacd_map_ms_mask_pns(ptr_aac_dec_channel_info);
        G.n_cpp_pns++;
      }
      G.n_cpp++;
      {
        int any_ms = 0, any_tns = 0;
        if (num_ch == 2) {
          for (int i = 0; i < 512 && !any_ms; i++) any_ms = ptr_aac_dec_channel_info[0]->common_window && rec[XAAC_SPS_MS_USED + i];
          for (int i = 0; i < 128 && !any_ms; i++) any_ms = ptr_aac_dec_channel_info[1]->ptr_code_book[i] >= 14;
        }
        for (int ch = 0; ch < num_ch; ch++) any_tns |= ptr_aac_dec_channel_info[ch]->str_tns_info.tns_data_present != 0;
        G.n_cpp_ms += any_ms;
        G.n_cpp_tns += any_tns;
      }
      return IA_NO_ERROR;
    }
  }
  G.n_cpp_ref++;
  return __real_ixheaacd_channel_pair_process(ptr_aac_dec_channel_info, num_ch, ptr_aac_tables, total_channels, object_type,
                                              aac_spect_data_resil_flag, aac_sf_data_resil_flag, in_data, out_data, self_ptr);
}

/* ================================ ixheaacd_dec_sbrdata ================================ */
IA_ERRORCODE __real_ixheaacd_dec_sbrdata(ia_sbr_header_data_struct *, ia_sbr_header_data_struct *, ia_sbr_frame_info_data_struct *,
                                         ia_sbr_prev_frame_data_struct *, ia_sbr_frame_info_data_struct *,
                                         ia_sbr_prev_frame_data_struct *, ixheaacd_misc_tables *, WORD32, WORD32, WORD32);
/* decoder/ixheaacd_env_dec.c:628 (called from ixheaacd_applysbr, decoder/ixheaacd_sbrdecoder.c:711, :1263): the fixed-point path
 * (xaacdec -esbr:0) runs on the GPU; USAC / enh_sbr frames keep the reference's float dequantisation (libm pow). */
IA_ERRORCODE __wrap_ixheaacd_dec_sbrdata(ia_sbr_header_data_struct *ptr_header_data_ch_0, ia_sbr_header_data_struct *ptr_header_data_ch_1,
                                         ia_sbr_frame_info_data_struct *ptr_sbr_data_ch_0,
                                         ia_sbr_prev_frame_data_struct *ptr_prev_data_ch_0,
                                         ia_sbr_frame_info_data_struct *ptr_sbr_data_ch_1,
                                         ia_sbr_prev_frame_data_struct *ptr_prev_data_ch_1, ixheaacd_misc_tables *ptr_common_tables,
                                         WORD32 ldmps_present, WORD32 audio_object_type, WORD32 ec_flag) {
  xaac_b200_ctx *c = b200_ctx();
  static int16_t rec[XAAC_SD_WORDS];
  if (c && !G.have_sd_rom) { /* the first call comes before the first ixheaacd_sbr_dec: the misc tables go to the device here */
    B200(xaac_b200_set_env_rom(c, &ixheaacd_aac_dec_env_calc_tables, 2404, ptr_common_tables, 2470), "set_env_rom");
    G.have_sd_rom = 1;
  }
  if (c &&
      b200_sd_pack(rec, ptr_header_data_ch_0, ptr_header_data_ch_1, ptr_sbr_data_ch_0, ptr_prev_data_ch_0, ptr_sbr_data_ch_1,
                   ptr_prev_data_ch_1, ldmps_present, audio_object_type, ec_flag) == 0) {
    const int was_err = ptr_header_data_ch_0->err_flag || (ptr_sbr_data_ch_1 && ptr_header_data_ch_1->err_flag);
    if (!G.d_sd) B200(xaac_b200_dev_alloc(c, sizeof(rec), (void **)&G.d_sd), "alloc");
    B200(xaac_b200_h2d(c, G.d_sd, rec, sizeof(rec)), "h2d sbrdata");
    B200(xaac_b200_dec_sbrdata_dev(c, G.d_sd, 1, NULL), "dec_sbrdata_dev");
    B200(xaac_b200_d2h(c, rec, G.d_sd, sizeof(rec)), "d2h sbrdata");
    b200_sd_unpack(rec, ptr_header_data_ch_0, ptr_header_data_ch_1, ptr_sbr_data_ch_0, ptr_prev_data_ch_0, ptr_sbr_data_ch_1,
                   ptr_prev_data_ch_1);
    G.n_sd++;
    G.n_sd_coupled += ptr_sbr_data_ch_1 && ptr_sbr_data_ch_0->coupling_mode;
    G.n_sd_concealed += !was_err && (ptr_header_data_ch_0->err_flag || (ptr_sbr_data_ch_1 && ptr_header_data_ch_1->err_flag));
    return rec[XAAC_SD_ERR] == 0 ? IA_NO_ERROR : (rec[XAAC_SD_ERR] == 2 ? (IA_ERRORCODE)-1 : (IA_ERRORCODE)IA_FATAL_ERROR);
  }
  G.n_sd_ref++;
  return __real_ixheaacd_dec_sbrdata(ptr_header_data_ch_0, ptr_header_data_ch_1, ptr_sbr_data_ch_0, ptr_prev_data_ch_0,
                                     ptr_sbr_data_ch_1, ptr_prev_data_ch_1, ptr_common_tables, ldmps_present, audio_object_type,
                                     ec_flag);
}

/* ================================ ixheaacd_decode_ps_data ================================ */
VOID __real_ixheaacd_decode_ps_data(ia_ps_dec_struct *ptr_ps_dec, WORD32 frame_size);
/* decoder/ixheaacd_ps_bitdec.c:98 (called from ixheaacd_applysbr, decoder/ixheaacd_sbrdecoder.c:723) */
VOID __wrap_ixheaacd_decode_ps_data(ia_ps_dec_struct *ptr_ps_dec, WORD32 frame_size) {
  xaac_b200_ctx *c = b200_ctx();
  static int16_t rec[XAAC_PSD_WORDS];
  if (!c) {
    __real_ixheaacd_decode_ps_data(ptr_ps_dec, frame_size);
    return;
  }
  b200_psd_pack(rec, ptr_ps_dec, frame_size);
  if (!G.d_psd) B200(xaac_b200_dev_alloc(c, sizeof(rec), (void **)&G.d_psd), "alloc");
  B200(xaac_b200_h2d(c, G.d_psd, rec, sizeof(rec)), "h2d psdata");
  B200(xaac_b200_decode_ps_data_dev(c, G.d_psd, 1, NULL), "decode_ps_data_dev");
  B200(xaac_b200_d2h(c, rec, G.d_psd, sizeof(rec)), "d2h psdata");
  b200_psd_unpack(rec, ptr_ps_dec);
  G.n_psd++;
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
