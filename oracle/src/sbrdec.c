/*
 * oracle/src/sbrdec.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain-C restatement of the fixed-point branch of ixheaacd_sbr_dec (decoder/ixheaacd_sbr_dec.c:662-1310 without
 * :816-1009) for the complex ("HQ") path, 1024-sample core frames (32 QMF slots, 6 overlap slots), non-ELD/LD object
 * types, no DRC, no MPS: overlap hand-over and ixheaacd_rescale_x_overlap, analysis, block-floating-point
 * bookkeeping, HF generation, envelope adjustment, LPC/overlap state update, [parametric stereo], synthesis.
 * It chains the stage oracles of qmf.c / hfgen.c / envcalc.c / ps.c.  Pinned against whole-stage records tapped from
 * real decodes of the compiled reference (tests/golden/sbrdec_tapped.npz).
 */
#include <string.h>
#include "fixmath.h"
#include "xaac_oracle.h"

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* decoder/ixheaacd_sbrdec_lpfuncs.c:453-527 (complex).  m = matrix rows 0..5 (the overlap slots). */
static void rescale_x_overlap(i32 *m, i16 *sf, i16 *misc, const i16 *env, int syn_usb) {
  int old_lsb = misc[XO_SBR_MISC_MAX_QMF_PREV];
  int start_slot = env[XO_ENV_TIME_STEP] * (misc[XO_SBR_MISC_END_POS_PREV] - env[XO_ENV_NUM_TIME_SLOTS]);
  int new_lsb = env[XO_ENV_MAX_QMF_SUBBAND];
  misc[XO_SBR_MISC_CODEC_USB] = (i16)new_lsb;
  misc[XO_SBR_MISC_SYN_LSB] = (i16)new_lsb;
  int b0 = imin(old_lsb, new_lsb), b1 = imax(old_lsb, new_lsb);
  if (new_lsb == old_lsb || old_lsb <= 0) return;
  for (int l = start_slot; l < 6; l++)
    for (int k = old_lsb; k < new_lsb; k++) m[128 * l + k] = m[128 * l + 64 + k] = 0;
  int source_scale, target_scale, t_lsb, t_usb;
  if (new_lsb > old_lsb) {
    source_scale = sf[XO_SF_OV_HB]; target_scale = sf[XO_SF_OV_LB]; t_lsb = 0; t_usb = old_lsb;
  } else {
    source_scale = sf[XO_SF_OV_LB]; target_scale = sf[XO_SF_OV_HB]; t_lsb = old_lsb; t_usb = syn_usb;
  }
  int reserve = xo_expsubbandsamples_hq(m, b0, b1, 0, start_slot);
  xo_adjust_scale_hq(m, b0, b1, 0, start_slot, reserve);
  source_scale += reserve;
  int delta = target_scale - source_scale;
  if (delta > 0) {
    delta = -delta;
    b0 = t_lsb;
    b1 = t_usb;
    if (new_lsb > old_lsb) sf[XO_SF_OV_LB] = (i16)source_scale;
    else sf[XO_SF_OV_HB] = (i16)source_scale;
  }
  xo_adjust_scale_hq(m, b0, b1, 0, start_slot, delta);
}

/* scratch: 40 rows x 128 WORD32 (2 LPC rows + 38 matrix rows).  time_in: 1024 samples at stride ch_in;
 * time_out: 2048 samples at stride ch_out (left); time_out_r likewise for the PS right channel (or NULL). */
int xo_sbr_dec_hq(const uint8_t *qrom, const uint8_t *env_rom, const uint8_t *misc_rom, const uint8_t *ps_rom,
                  const i16 *side, i16 *st, i16 *ps_st, const i16 *time_in, int ch_in, i16 *time_out, i16 *time_out_r,
                  int ch_out, i32 *scratch) {
  const i16 *env = side + XO_SIDE_ENV;
  const i16 *hfs = side + XO_SIDE_HF;
  const int apply = side[XO_SIDE_APPLY], ps_on = side[XO_SIDE_PS];
  i16 *sf = st + XO_SBR_ST_SF, *misc = st + XO_SBR_ST_MISC;
  i32 *lpc = (i32 *)(st + XO_SBR_ST_LPC), *ov = (i32 *)(st + XO_SBR_ST_OV), *bw_prev = (i32 *)(st + XO_SBR_ST_BW_PREV);
  i32 *m = scratch + 256;
  const i16 *border = env + XO_ENV_BORDER_VEC;
  const int num_env = env[XO_ENV_NUM_ENV];

  memcpy(m, ov, 6 * 128 * sizeof(i32));               /* sbr_dec.c:749-753 */
  sf[XO_SF_LB] = 0;                                   /* :768 */
  if (apply) rescale_x_overlap(m, sf, misc, env, misc[XO_SBR_MISC_SYN_USB]);

  { /* :1025-1029 */
    i32 pos = st[XO_SBR_ST_ANAL_POS], fpos = st[XO_SBR_ST_ANAL_POS + 1];
    sf[XO_SF_ST_LB] = 0; /* generic:630 */
    sf[XO_SF_LB] = (i16)xo_anal_qmffilt_hq(qrom, time_in, ch_in, st + XO_SBR_ST_ANAL_STATES, &pos, &fpos,
                                           misc[XO_SBR_MISC_CODEC_USB], m + 6 * 128);
    st[XO_SBR_ST_ANAL_POS] = (i16)pos;
    st[XO_SBR_ST_ANAL_POS + 1] = (i16)fpos;
  }
  int save_lb_scale;
  { /* :1050-1114 */
    int usb = misc[XO_SBR_MISC_CODEC_USB];
    int reserve = xo_expsubbandsamples_hq(m, 0, usb, 6, 38);
    int reserve_ov1 = xo_expsubbandsamples_hq(m, 0, usb, 0, 6);
    int reserve_ov2 = xo_expsubbandsamples_hq(lpc, 0, usb, 0, 2);
    reserve_ov1 = imin(reserve_ov1, reserve_ov2);
    int shift1 = sf[XO_SF_LB] + reserve, shift2 = sf[XO_SF_OV_LB] + reserve_ov1;
    int min_shift = imin(shift1, shift2);
    int shift_over = shift2 - min_shift;
    reserve -= shift1 - min_shift;
    sf[XO_SF_OV_LB] = (i16)(sf[XO_SF_OV_LB] + (reserve_ov1 - shift_over));
    xo_adjust_scale_hq(m, 0, usb, 0, 6, reserve_ov1 - shift_over);
    xo_adjust_scale_hq(m, 0, usb, 6, 38, reserve);
    xo_adjust_scale_hq(lpc, 0, usb, 0, 2, reserve_ov1 - shift_over);
    sf[XO_SF_LB] = (i16)(sf[XO_SF_LB] + reserve);
    save_lb_scale = sf[XO_SF_LB];
  }
  for (int l = 6; l < 38; l++) { /* :1117-1127 */
    memset(m + 128 * l + 32, 0, 32 * sizeof(i32));
    memset(m + 128 * l + 64 + 32, 0, 32 * sizeof(i32));
  }
  if (apply) {
    i16 hf[XO_HF_PRM_WORDS];
    memcpy(hf, hfs, sizeof(hf));
    hf[XO_HF_FACTOR] = env[XO_ENV_TIME_STEP];
    hf[XO_HF_START_IDX] = border[0];
    hf[XO_HF_STOP_IDX] = ox_sub16_sat(border[num_env], env[XO_ENV_NUM_TIME_SLOTS]);
    for (int i = 0; i < 10; i++) hf[XO_HF_INVF_PREV + i] = misc[XO_SBR_MISC_INVF_PREV + i];
    hf[XO_HF_OV_LB_SCALE] = sf[XO_SF_OV_LB];
    hf[XO_HF_LB_SCALE] = sf[XO_SF_LB];
    hf[XO_HF_MAX_QMF_SUBBAND] = env[XO_ENV_MAX_QMF_SUBBAND];
    sf[XO_SF_HB] = (i16)xo_hf_generator_hq(lpc, m, hf, bw_prev);                      /* :1169-1180 */
    i16 envp[XO_ENV_PRM_WORDS];
    memcpy(envp, env, sizeof(envp));
    envp[XO_ENV_MAX_QMF_SUBBAND_PREV] = misc[XO_SBR_MISC_MAX_QMF_PREV];
    int err = xo_calc_sbrenvelope_hq(env_rom, misc_rom, envp, sf, st + XO_SBR_ST_ENV, m); /* :1195-1202 */
    if (err) return err;
    for (int i = 0; i < hf[XO_HF_NUM_IF_BANDS]; i++) misc[XO_SBR_MISC_INVF_PREV + i] = hf[XO_HF_INVF + i]; /* :1205-1213 */
    misc[XO_SBR_MISC_MAX_QMF_PREV] = env[XO_ENV_MAX_QMF_SUBBAND];
    misc[XO_SBR_MISC_END_POS_PREV] = border[num_env];
  } else {
    sf[XO_SF_HB] = (i16)save_lb_scale;
  }
  { /* :1218-1245 */
    int usb = misc[XO_SBR_MISC_CODEC_USB];
    for (int i = 0; i < 2; i++) {
      memcpy(lpc + 128 * i, m + 128 * (30 + i), usb * sizeof(i32));
      memcpy(lpc + 128 * i + 64, m + 128 * (30 + i) + 64, usb * sizeof(i32));
    }
  }
  /* :1284-1290 — the reference copies NO_SYNTHESIS_CHANNELS * op_delay = 384 words *before* doubling the count for
   * the complex layout, i.e. only slots 32..34 (re|im) reach the overlap buffer; slots 3..5 of the buffer keep their
   * previous content.  Rows 32..37 are not touched by the synthesis, so the copy can be taken here. */
  i32 ovsave[3 * 128];
  memcpy(ovsave, m + 32 * 128, sizeof(ovsave));

  i32 sfv[4];
  i32 off = st[XO_SBR_ST_SYN_POS], fpos = st[XO_SBR_ST_SYN_POS + 1];
  if (apply && ps_on && env[XO_ENV_CHANNEL_MODE] == 3) {
    xo_ps_synth_pair(qrom, env_rom, misc_rom, ps_rom, side + XO_SIDE_PS_PRM, st, ps_st, m, time_out, time_out_r,
                     ch_out, ps_on != 2); /* :1247-1272 */
  } else {
    sfv[0] = sf[XO_SF_OV_LB]; sfv[1] = sf[XO_SF_LB]; sfv[2] = sf[XO_SF_HB]; sfv[3] = sf[XO_SF_ST_SYN];
    xo_synt_qmffilt_hq(qrom, m, st + XO_SBR_ST_SYN_STATES, &off, &fpos, sfv, misc[XO_SBR_MISC_SYN_LSB],
                       misc[XO_SBR_MISC_SYN_USB], 6, time_out, ch_out); /* :1273-1281 */
    st[XO_SBR_ST_SYN_POS] = (i16)off;
    st[XO_SBR_ST_SYN_POS + 1] = (i16)fpos;
  }
  memcpy(ov, ovsave, sizeof(ovsave));
  sf[XO_SF_OV_LB] = (i16)save_lb_scale; /* :1308 */
  return 0;
}

/* Stage glue.  mode 0: the SBR hand-over of ixheaacd_allocate_sbr_scr (decoder/ixheaacd_api.c:337-370),
 * round16(shl32_sat(x, qshift_adj)); mode 1: the AAC-LC output stage with the peak limiter off — ixheaacd_scale_adjust
 * (decoder/ixheaacd_peak_limiter.c:324-333, wrapping x * (1 << qshift_adj)) followed by round16
 * (decoder/ixheaacd_api.c:3676-3681). */
void xo_imdct_out_to_pcm16(const i32 *in, const int8_t *qshift_adj, i16 *out, int n_units, int mode) {
  for (int u = 0; u < n_units; u++)
    for (int i = 0; i < 1024; i++) {
      i32 x = in[(size_t)u * 1024 + i];
      int q = qshift_adj[u];
      x = mode ? (i32)((u32)x * (u32)(1 << q)) : ox_shl32_sat(x, q);
      out[(size_t)u * 1024 + i] = ox_round16(x);
    }
}
