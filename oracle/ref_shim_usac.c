/*
 * oracle/ref_shim_usac.c — TEST INFRASTRUCTURE ONLY.
 *
 * Flat-array entry points around the UNMODIFIED reference for the USAC frequency-domain core path:
 * ixheaacd_fd_frm_dec (decoder/ixheaacd_imdct.c:596) and ixheaacd_complex_fft (decoder/ixheaacd_fft.c:2664).
 * Compiled against the reference headers where they lie and linked into oracle/_ref/libxaac_ref.so (oracle/Makefile).
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "ixheaac_type_def.h"
#include "ixheaacd_interface.h"
#include "ixheaacd_defines.h"
#include "ixheaacd_aac_rom.h"
#include "ixheaacd_bitbuffer.h"
#include "ixheaacd_tns_usac.h"
#include "ixheaacd_cnst.h"
#include "ixheaacd_acelp_info.h"
#include "ixheaacd_td_mdct.h"
#include "ixheaacd_sbrdecsettings.h"
#include "ixheaacd_info.h"
#include "ixheaacd_sbr_common.h"
#include "ixheaacd_drc_data_struct.h"
#include "ixheaacd_drc_dec.h"
#include "ixheaacd_sbrdecoder.h"
#include "ixheaacd_mps_polyphase.h"
#include "ixheaac_sbr_const.h"
#include "ixheaacd_pulsedata.h"
#include "ixheaacd_pns.h"
#include "ixheaacd_lt_predict.h"
#include "ixheaacd_ec_defines.h"
#include "ixheaacd_ec_struct_def.h"
#include "ixheaacd_main.h"
#include "ixheaacd_windows.h"

extern const WORD32 ixheaacd_twiddle_table_fft_32x32[514];
extern const WORD32 ixheaacd_pre_post_twid_cos_512[512];
extern const WORD32 ixheaacd_pre_post_twid_sin_512[512];
extern const WORD32 ixheaacd_pre_post_twid_cos_64[64];
extern const WORD32 ixheaacd_pre_post_twid_sin_64[64];
VOID ixheaacd_complex_fft(WORD32 *data_r, WORD32 *data_i, WORD32 nlength, WORD32 fft_mode, WORD32 *preshift);

/* ROM blob of the USAC FD path, layout = XO_UROM_* in oracle/src/xaac_oracle.h (15880 bytes) */
const void *ref_rom_usac_tables(int *bytes) {
  static int32_t blob[3970];
  int32_t *p = blob;
  memcpy(p, ixheaacd_twiddle_table_fft_32x32, 514 * 4); p += 514;
  memcpy(p, ixheaacd_pre_post_twid_cos_512, 512 * 4); p += 512;
  memcpy(p, ixheaacd_pre_post_twid_sin_512, 512 * 4); p += 512;
  memcpy(p, ixheaacd_pre_post_twid_cos_64, 64 * 4); p += 64;
  memcpy(p, ixheaacd_pre_post_twid_sin_64, 64 * 4); p += 64;
  memcpy(p, ixheaacd_sine_win_1024, 1024 * 4); p += 1024;
  memcpy(p, ixheaacd_kbd_win1024, 1024 * 4); p += 1024;
  memcpy(p, ixheaacd_sine_win_128, 128 * 4); p += 128;
  memcpy(p, ixheaacd_kbd_win128, 128 * 4); p += 128;
  if (bytes) *bytes = (int)((p - blob) * 4);
  return blob;
}

/* ixheaacd_complex_fft (fft_mode = 1, the IMDCT's use): xr, xi [n] in/out; returns the updated *preshift */
int ref_usac_complex_fft(int32_t *xr, int32_t *xi, int n, int preshift) {
  WORD32 ps = preshift;
  ixheaacd_complex_fft(xr, xi, n, 1, &ps);
  return ps;
}

/* ixheaacd_fd_frm_dec for one channel of a pure frequency-domain stream (previous frame FD, no FAC data, frame ok):
 *   coef [1024] dequantised spectrum (destroyed), overlap [1024] in/out, out [1024] WORD32 time samples (Q15 scale).
 * The caller carries window_shape_prev exactly as ixheaacd_core_coder_data does (ext_ch_ele.c:968). */
int ref_usac_fd_frm_dec(int32_t *coef, int32_t *overlap, int win_seq, int win_shape, int win_shape_prev, int32_t *out) {
  static __thread ia_usac_data_struct *d;
  if (!d) d = (ia_usac_data_struct *)calloc(1, sizeof(ia_usac_data_struct));
  d->ccfl = 1024;
  d->coef_fix[0] = coef;
  memcpy(d->overlap_data_ptr[0], overlap, 1024 * sizeof(int32_t));
  d->window_shape[0] = win_shape;
  d->window_shape_prev[0] = win_shape_prev;
  d->window_sequence[0] = win_seq;
  d->td_frame_prev[0] = 0;
  d->fac_data_present[0] = 0;
  d->ec_flag = 0;
  d->frame_ok = 1;
  d->str_tddec[0] = NULL;
  int err = (int)ixheaacd_fd_frm_dec(d, 0);
  memcpy(overlap, d->overlap_data_ptr[0], 1024 * sizeof(int32_t));
  memcpy(out, d->output_data_ptr[0], 1024 * sizeof(int32_t));
  return err;
}
void ref_usac_fd_frm_dec_batch(int32_t *coef, int32_t *overlap, const int32_t *win_seq, const int32_t *win_shape,
                               const int32_t *win_shape_prev, int32_t *out, int32_t *err, int n) {
  for (int u = 0; u < n; u++)
    err[u] = ref_usac_fd_frm_dec(coef + (size_t)u * 1024, overlap + (size_t)u * 1024, win_seq[u], win_shape[u],
                                 win_shape_prev[u], out + (size_t)u * 1024);
}

/* ------------------------------------------------------------------------------------------------
 * AAC-LC output stage: ixheaacd_peak_limiter_init / ixheaacd_peak_limiter_process
 * (decoder/ixheaacd_peak_limiter.c:45-75, 177-307) driven from the flat XO_PL_* record of oracle/src/xaac_oracle.h.
 * ---------------------------------------------------------------------------------------------- */
#include "ixheaacd_peak_limiter_struct_def.h"
#include "ixheaac_constants.h"
#include "ixheaac_basic_ops32.h"
#include "ixheaac_basic_ops16.h"
#include "src/xaac_oracle.h"
WORD32 ixheaacd_peak_limiter_init(ia_peak_limiter_struct *, UWORD32, UWORD32, FLOAT32 *, UWORD32 *);
VOID ixheaacd_peak_limiter_process(ia_peak_limiter_struct *, VOID *, UWORD32, UWORD8 *);

static void pl_pack(int32_t *st, const ia_peak_limiter_struct *p) {
  memset(st, 0, XO_PL_WORDS * 4);
  memcpy(st + XO_PL_ATTACK_CONST, &p->attack_constant, 4);
  memcpy(st + XO_PL_RELEASE_CONST, &p->release_constant, 4);
  memcpy(st + XO_PL_GAIN_MOD, &p->gain_modified, 4);
  memcpy(st + XO_PL_MIN_GAIN, &p->min_gain, 4);
  memcpy(st + XO_PL_PSG, &p->pre_smoothed_gain, 8);
  st[XO_PL_ATTACK] = (int32_t)p->attack_time_samples;
  st[XO_PL_DELAY_IDX] = (int32_t)p->delayed_input_index;
  st[XO_PL_MAX_IDX] = p->max_idx;
  st[XO_PL_CIR] = p->cir_buf_pnt;
  st[XO_PL_LIMITER_ON] = (int32_t)p->limiter_on;
  st[XO_PL_NUM_CH] = (int32_t)p->num_channels;
  memcpy(st + XO_PL_MAX_BUF, p->max_buf, 4 * p->attack_time_samples);
  memcpy(st + XO_PL_DELAYED, p->delayed_input, 4 * p->attack_time_samples * p->num_channels);
}
static void pl_unpack(ia_peak_limiter_struct *p, const int32_t *st) {
  memset(p, 0, sizeof(*p));
  memcpy(&p->attack_constant, st + XO_PL_ATTACK_CONST, 4);
  memcpy(&p->release_constant, st + XO_PL_RELEASE_CONST, 4);
  memcpy(&p->gain_modified, st + XO_PL_GAIN_MOD, 4);
  memcpy(&p->min_gain, st + XO_PL_MIN_GAIN, 4);
  memcpy(&p->pre_smoothed_gain, st + XO_PL_PSG, 8);
  p->attack_time_samples = (UWORD32)st[XO_PL_ATTACK];
  p->delayed_input_index = (UWORD32)st[XO_PL_DELAY_IDX];
  p->max_idx = st[XO_PL_MAX_IDX];
  p->cir_buf_pnt = st[XO_PL_CIR];
  p->limiter_on = (UWORD32)st[XO_PL_LIMITER_ON];
  p->num_channels = (UWORD32)st[XO_PL_NUM_CH];
  p->max_buf = p->buffer;
  p->delayed_input = p->buffer + p->attack_time_samples * 4 + 32;
  memcpy(p->max_buf, st + XO_PL_MAX_BUF, 4 * p->attack_time_samples);
  memcpy(p->delayed_input, st + XO_PL_DELAYED, 4 * p->attack_time_samples * p->num_channels);
}
/* state of a freshly initialised limiter (ixheaacd_peak_limiter_init); returns delay_in_samples */
int ref_peak_limiter_init(int32_t *st, int num_channels, int sample_rate) {
  static __thread ia_peak_limiter_struct p;
  UWORD32 delay = 0;
  memset(&p, 0, sizeof(p));
  ixheaacd_peak_limiter_init(&p, (UWORD32)num_channels, (UWORD32)sample_rate, p.buffer, &delay);
  pl_pack(st, &p);
  return (int)delay;
}
/* one frame: samples interleaved WORD32 [1024][ch] in place; pcm16 = round16 of the result (api.c:3676-3681) */
void ref_peak_limiter_batch(int32_t *st, int32_t *samples, const int8_t *qshift_adj, int16_t *pcm16, int ch, int n) {
  static __thread ia_peak_limiter_struct p;
  for (int u = 0; u < n; u++) {
    int32_t *s = st + (size_t)u * XO_PL_WORDS, *x = samples + (size_t)u * 1024 * ch;
    pl_unpack(&p, s);
    ixheaacd_peak_limiter_process(&p, x, 1024, (UWORD8 *)(qshift_adj + (size_t)u * ch));
    pl_pack(s, &p);
    if (pcm16)
      for (int i = 0; i < 1024 * ch; i++) pcm16[(size_t)u * 1024 * ch + i] = ixheaac_round16(x[i]);
  }
}
