/*
 * oracle/ref_shim_hbe.c — TEST INFRASTRUCTURE ONLY.
 *
 * Flat-array entry points around the UNMODIFIED reference's QMF harmonic transposer ixheaacd_qmf_hbe_apply
 * (decoder/ixheaacd_hbe_trans.c:224) in the XO_HBE_* layouts of oracle/src/xaac_oracle.h, and the tap installed with
 * `ld --wrap=ixheaacd_qmf_hbe_apply` into oracle/_ref/xaacdec_tap.  Compiled against the reference headers where they lie.
 *   <tap>.hbe record: int32 'HBE1', int32 ret, int32 cfg[16], float state_in[3616], float qmf_re[2048], qmf_im[2048],
 *                     float pv_re[2048], pv_im[2048] (after the call), float state_out[3616]
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
#include "ixheaacd_sbrqmftrans.h"
#include "ixheaacd_qmf_poly.h"
#include "ixheaac_esbr_rom.h"
#include "src/xaac_oracle.h"

extern const FLOAT32 ixheaac_twiddle_table_fft_float[514];
extern const FLOAT32 ixheaac_twidle_tbl_48[64];
extern const FLOAT32 ixheaac_twidle_tbl_24[32];
VOID ixheaacd_esbr_hbe_data_init(ia_esbr_hbe_txposer_struct *pstr_esbr_hbe_txposer, const WORD32 num_aac_samples,
                                 WORD32 samp_fac_4_flag, const WORD32 num_out_samples, VOID *persistent_hbe_mem,
                                 WORD32 *total_persistant);

/* ROM blob of the transposer, layout XO_HROM_* */
const void *ref_rom_hbe_tables(int *bytes) {
  static float blob[XO_HROM_WORDS];
  memset(blob, 0, sizeof(blob));
  memcpy(blob + XO_HROM_WIN, ixheaac_sub_samp_qmf_window_coeff, 1560 * 4);
  memcpy(blob + XO_HROM_SYNCOS, ixheaac_synth_cos_table_kl_4, 16 * 4);
  memcpy(blob + XO_HROM_SYNCOS + 16, ixheaac_synth_cos_table_kl_8, 32 * 4);
  memcpy(blob + XO_HROM_SYNCOS + 48, ixheaac_synth_cos_table_kl_12, 48 * 4);
  memcpy(blob + XO_HROM_SYNCOS + 96, ixheaac_synth_cos_table_kl_16, 64 * 4);
  memcpy(blob + XO_HROM_ANACS, ixheaac_analy_cos_sin_table_kl_8, 32 * 4);
  memcpy(blob + XO_HROM_ANACS + 32, ixheaac_analy_cos_sin_table_kl_16, 64 * 4);
  memcpy(blob + XO_HROM_ANACS + 96, ixheaac_analy_cos_sin_table_kl_24, 96 * 4);
  memcpy(blob + XO_HROM_ANACS + 192, ixheaac_analy_cos_sin_table_kl_32, 128 * 4);
  memcpy(blob + XO_HROM_COSTRANS, ixheaac_cos_table_trans_qmf, 448 * 4);
  memcpy(blob + XO_HROM_FFTTW, ixheaac_twiddle_table_fft_float, 514 * 4);
  memcpy(blob + XO_HROM_TW24, ixheaac_twidle_tbl_24, 32 * 4);
  memcpy(blob + XO_HROM_TW48, ixheaac_twidle_tbl_48, 64 * 4);
  memcpy(blob + XO_HROM_PVCOS, ixheaac_phase_vocoder_cos_table, 64 * 4);
  memcpy(blob + XO_HROM_PVSIN, ixheaac_phase_vocoder_sin_table, 64 * 4);
  memcpy(blob + XO_HROM_INTERP, ixheaac_hbe_post_anal_proc_interp_coeff, 8 * 4);
  memcpy(blob + XO_HROM_SELCASE, ixheaac_sel_case, 40 * 4);
  memcpy(blob + XO_HROM_XP2, ixheaac_hbe_x_prod_cos_table_trans_2, 512 * 4);
  memcpy(blob + XO_HROM_XP3, ixheaac_hbe_x_prod_cos_table_trans_3, 512 * 4);
  memcpy(blob + XO_HROM_XP4, ixheaac_hbe_x_prod_cos_table_trans_4, 512 * 4);
  memcpy(blob + XO_HROM_XP41, ixheaac_hbe_x_prod_cos_table_trans_4_1, 512 * 4);
  memcpy(blob + XO_HROM_SYN20, ixheaac_synth_cos_table_kl_20, 800 * 4);
  memcpy(blob + XO_HROM_ANA40, ixheaac_analy_cos_sin_table_kl_40, 3200 * 4);
  if (bytes) *bytes = (int)sizeof(blob);
  return blob;
}

/* ---- flat <-> ia_esbr_hbe_txposer_struct ---- */
void hbe_cfg_pack(int32_t *cfg, const ia_esbr_hbe_txposer_struct *t, int pitch) {
  memset(cfg, 0, 4 * XO_HBE_CFG_WORDS);
  cfg[XO_HBE_SYNTH_SIZE] = t->synth_size;
  cfg[XO_HBE_K_START] = t->k_start;
  cfg[XO_HBE_START_BAND] = t->start_band;
  cfg[XO_HBE_END_BAND] = t->end_band;
  cfg[XO_HBE_MAX_STRETCH] = t->max_stretch;
  cfg[XO_HBE_PITCH] = pitch;
  cfg[XO_HBE_USF4] = t->upsamp_4_flag;
  for (int i = 0; i < 6; i++) cfg[XO_HBE_XOVER + i] = t->x_over_qmf[i];
}
/* returns 0 when the instance satisfies the invariants the flat state relies on */
int hbe_state_pack(float *st, const ia_esbr_hbe_txposer_struct *t) {
  const int S = t->synth_size;
  int bad = 0;
  memset(st, 0, 4 * XO_HBE_ST_WORDS);
  if (S < 1 || S > 20 || t->no_bins != 32) return -1;
  memcpy(st + XO_HBE_ST_TAIL, t->ptr_input_buf + t->no_bins * S, S * 4);
  memcpy(st + XO_HBE_ST_SYNTH, t->synth_buf, 18 * S * 4);
  memcpy(st + XO_HBE_ST_ANAL, t->analy_buf, 18 * S * 4);
  for (int r = 0; r < 12; r++) memcpy(st + XO_HBE_ST_QIN + 128 * r, t->qmf_in_buf[16 + r], 512);
  for (int r = 0; r < 10; r++) memcpy(st + XO_HBE_ST_QOUT + 128 * r, t->qmf_out_buf[32 + r], 512);
  for (int r = 42; r < 64; r++)
    for (int c = 0; c < 128; c++)
      if (t->qmf_out_buf[r][c] != 0.0f) bad = 1;
  return bad;
}
static void hbe_state_unpack(ia_esbr_hbe_txposer_struct *t, const float *st) {
  const int S = t->synth_size;
  memcpy(t->ptr_input_buf + t->no_bins * S, st + XO_HBE_ST_TAIL, S * 4);
  memcpy(t->synth_buf, st + XO_HBE_ST_SYNTH, 18 * S * 4);
  memcpy(t->analy_buf, st + XO_HBE_ST_ANAL, 18 * S * 4);
  for (int r = 0; r < 12; r++) memcpy(t->qmf_in_buf[16 + r], st + XO_HBE_ST_QIN + 128 * r, 512);
  for (int r = 0; r < 10; r++) memcpy(t->qmf_out_buf[32 + r], st + XO_HBE_ST_QOUT + 128 * r, 512);
}

static const FLOAT32 *prot(int len) { /* = the file-static ixheaacd_map_prot_filter (hbe_trans.c:70-101) */
  static const int off[] = {0, 40, 120, 240, 400, 600, 840, 1160};
  static const int lens[] = {4, 8, 12, 16, 20, 24, 32, 40};
  for (int i = 0; i < 8; i++)
    if (lens[i] == len) return &ixheaac_sub_samp_qmf_window_coeff[off[i]];
  return &ixheaac_sub_samp_qmf_window_coeff[0];
}

/* drive the compiled ixheaacd_qmf_hbe_apply from flat records.  tbl = {num_lo, num_hi, freq_band_table[LOW][0..num_lo],
 * freq_band_table[HIGH][0..num_hi]} or NULL: the tables the reference re-initialises from when its FFT pointers are NULL
 * (synth_size 20); they must reproduce cfg (use ref_esbr_hbe_reinit to derive cfg from them). */
int ref_esbr_hbe_apply_tbl(const int32_t *cfg, float *state, const float *qmf_re, const float *qmf_im, float *pv_re,
                           float *pv_im, const int16_t *tbl) {
  static __thread ia_esbr_hbe_txposer_struct t;
  static __thread double pers[16384]; /* 128 KB >= MAX_HBE_PERSISTENT_SIZE for the 2:1 system */
  static __thread ia_sbr_header_data_struct hd;
  WORD32 used = 0;
  memset(pers, 0, sizeof(pers));
  ixheaacd_esbr_hbe_data_init(&t, 1024, 0, 2048, pers, &used);
  if ((size_t)used > sizeof(pers)) return -100;
  const int S = cfg[XO_HBE_SYNTH_SIZE];
  t.synth_size = S;
  t.k_start = cfg[XO_HBE_K_START];
  t.start_band = cfg[XO_HBE_START_BAND];
  t.end_band = cfg[XO_HBE_END_BAND];
  t.max_stretch = cfg[XO_HBE_MAX_STRETCH];
  t.upsamp_4_flag = cfg[XO_HBE_USF4];
  t.esbr_hq = 0;
  t.synth_buf_offset = 18 * S;
  for (int i = 0; i < 6; i++) t.x_over_qmf[i] = cfg[XO_HBE_XOVER + i];
  switch (S) { /* hbe_trans.c:127-177 */
    case 4:
      t.synth_cos_tab = (FLOAT32 *)ixheaac_synth_cos_table_kl_4;
      t.analy_cos_sin_tab = (FLOAT32 *)ixheaac_analy_cos_sin_table_kl_8;
      t.ixheaacd_real_synth_fft = &ixheaac_real_synth_fft_p2;
      t.ixheaacd_cmplx_anal_fft = &ixheaac_cmplx_anal_fft_p2;
      break;
    case 8:
      t.synth_cos_tab = (FLOAT32 *)ixheaac_synth_cos_table_kl_8;
      t.analy_cos_sin_tab = (FLOAT32 *)ixheaac_analy_cos_sin_table_kl_16;
      t.ixheaacd_real_synth_fft = &ixheaac_real_synth_fft_p2;
      t.ixheaacd_cmplx_anal_fft = &ixheaac_cmplx_anal_fft_p2;
      break;
    case 12:
      t.synth_cos_tab = (FLOAT32 *)ixheaac_synth_cos_table_kl_12;
      t.analy_cos_sin_tab = (FLOAT32 *)ixheaac_analy_cos_sin_table_kl_24;
      t.ixheaacd_real_synth_fft = &ixheaac_real_synth_fft_p3;
      t.ixheaacd_cmplx_anal_fft = &ixheaac_cmplx_anal_fft_p3;
      break;
    case 16:
      t.synth_cos_tab = (FLOAT32 *)ixheaac_synth_cos_table_kl_16;
      t.analy_cos_sin_tab = (FLOAT32 *)ixheaac_analy_cos_sin_table_kl_32;
      t.ixheaacd_real_synth_fft = &ixheaac_real_synth_fft_p2;
      t.ixheaacd_cmplx_anal_fft = &ixheaac_cmplx_anal_fft_p2;
      break;
    case 20: /* no FFT pointers: the reference re-initialises inside every call (needs the frequency tables: not drivable
              * from the flat cfg unless they reproduce it, so the shim hands it tables that do) */
      t.synth_cos_tab = (FLOAT32 *)ixheaac_synth_cos_table_kl_20;
      t.analy_cos_sin_tab = (FLOAT32 *)ixheaac_analy_cos_sin_table_kl_40;
      break;
    default:
      return -2;
  }
  t.synth_wind_coeff = (FLOAT32 *)prot(S);
  t.analy_wind_coeff = (FLOAT32 *)prot(2 * S);
  hbe_state_unpack(&t, state);
  memset(&hd, 0, sizeof(hd));
  static __thread ia_freq_band_data_struct fb;
  static __thread WORD16 lo[64], hi[64];
  if (tbl) {
    memset(&fb, 0, sizeof(fb));
    fb.num_sf_bands[0] = tbl[0];
    fb.num_sf_bands[1] = tbl[1];
    memcpy(lo, tbl + 2, 2 * (tbl[0] + 1));
    memcpy(hi, tbl + 2 + tbl[0] + 1, 2 * (tbl[1] + 1));
    fb.freq_band_table[0] = lo;
    fb.freq_band_table[1] = hi;
    hd.pstr_freq_band_data = &fb;
  } else if (S == 20) {
    return -2;
  }
  static __thread FLOAT32 in_re[32][64], in_im[32][64], o_re[32][64], o_im[32][64];
  memcpy(in_re, qmf_re, sizeof(in_re));
  memcpy(in_im, qmf_im, sizeof(in_im));
  memcpy(o_re, pv_re, sizeof(o_re));
  memcpy(o_im, pv_im, sizeof(o_im));
  WORD32 ret = ixheaacd_qmf_hbe_apply(&t, in_re, in_im, 32, o_re, o_im, cfg[XO_HBE_PITCH], &hd);
  memcpy(pv_re, o_re, sizeof(o_re));
  memcpy(pv_im, o_im, sizeof(o_im));
  if (ret == 0 && hbe_state_pack(state, &t) != 0) return -101;
  if (ret == 0 && tbl) { /* the re-initialised instance must be the one cfg describes */
    int32_t c2[XO_HBE_CFG_WORDS];
    hbe_cfg_pack(c2, &t, cfg[XO_HBE_PITCH]);
    for (int i = 0; i < XO_HBE_CFG_WORDS; i++)
      if (i != XO_HBE_REINIT && c2[i] != cfg[i]) return -102;
  }
  return ret;
}
/* tbl: [n][128] int16 rows in the layout above, or NULL */
void ref_esbr_hbe_apply_batch(const int32_t *cfg, float *state, const float *qmf_re, const float *qmf_im, float *pv_re,
                              float *pv_im, const int16_t *tbl, int32_t *err, int n) {
  for (int u = 0; u < n; u++) {
    int e = ref_esbr_hbe_apply_tbl(cfg + (size_t)u * XO_HBE_CFG_WORDS, state + (size_t)u * XO_HBE_ST_WORDS,
                                   qmf_re + (size_t)u * 2048, qmf_im + (size_t)u * 2048, pv_re + (size_t)u * 2048,
                                   pv_im + (size_t)u * 2048, tbl ? tbl + (size_t)u * 128 : NULL);
    if (err) err[u] = e;
  }
}

#ifndef XAAC_REF_TAPS
/* ---- CPU baseline of bench.py --workload xheaac_stereo_chain (BASELINE configs[4]): one stereo xHE-AAC frame per pair of
 * units through the reference's own functions, harmonic transposer included: ixheaacd_fd_frm_dec -> x 2^-15 -> the eSBR branch
 * of ixheaacd_sbr_dec with hbe_flag = 1 spelled out with its leaf functions (history memmoves incl. the 32-slot codec delay,
 * ixheaacd_esbr_analysis_filt_block, ixheaacd_qmf_hbe_apply, ixheaacd_generate_hf, ixheaacd_sbr_env_calc, regrouping, synthesis
 * core) -> ixheaacd_samples_sat.  q6 = [n] x {qmf_re[72*64], qmf_im[72*64], out_re[2560], out_im[2560], pv_re[2560], pv_im[2560]}.
 * The transposer instance is rebuilt from the flat state on every call (about 130 KB of memset / memcpy per unit in this arm
 * that the real decoder does not pay: a few per cent of the unit's time). */
int ref_usac_fd_frm_dec(int32_t *coef, int32_t *overlap, int win_seq, int win_shape, int win_shape_prev, int32_t *out);
void ref_esbr_anal32(const float *time_in, int32_t *states, int32_t *pos, float *qmf);
int ref_esbr_generate_hf(const float *src_re, const float *src_im, const float *pv_re, const float *pv_im, float *dst_re,
                         float *dst_im, const int32_t *par, float *bw_prev, int32_t *patch_out);
int ref_esbr_env_calc(float *re, float *im, int32_t *ipar, const float *fpar, float *state);
void ref_esbr_synth64(const float *qmf, int32_t *fs, int32_t *pos, float *out);
#define Q6_WORDS (2 * 4608 + 4 * 2560)
void ref_xheaac_hbe_chain_batch(int32_t *coef, int32_t *overlap, const int32_t *win_seq, const int32_t *win_shape,
                                const int32_t *win_shape_prev, float *q6, int32_t *anal, int32_t *apos, int32_t *synth,
                                int32_t *spos, float *bw, int32_t *patch, float *ec, float *hbe_state, const int32_t *hbe_cfg,
                                const int16_t *hbe_tbl, const int32_t *hf_par, int32_t *ec_ipar, const float *ec_fpar,
                                const int32_t *rg, int16_t *pcm /* [n/2][2048][2] */, int32_t *err, int a, int b) {
  static __thread int32_t core[1024];
  static __thread float tin[1024], qa[32 * 128], m[32 * 128], tout[2048], lre[2560], lim[2560];
  for (int u = a; u < b; u++) {
    float *q = q6 + (size_t)u * Q6_WORDS;
    float *qre = q, *qim = q + 4608, *ore = q + 9216, *oim = ore + 2560, *pre = oim + 2560, *pim = pre + 2560;
    int e = ref_usac_fd_frm_dec(coef + (size_t)u * 1024, overlap + (size_t)u * 1024, win_seq[u], win_shape[u],
                                win_shape_prev[u], core);
    for (int k = 0; k < 1024; k++) tin[k] = (FLOAT32)((FLOAT32)core[k] * (FLOAT32)(0.000030517578125));
    memmove(qre, qre + 32 * 64, 40 * 64 * sizeof(float));
    memmove(qim, qim + 32 * 64, 40 * 64 * sizeof(float));
    memmove(ore, ore + 32 * 64, 8 * 64 * sizeof(float));
    memmove(oim, oim + 32 * 64, 8 * 64 * sizeof(float));
    memmove(pre, pre + 32 * 64, 8 * 64 * sizeof(float));
    memmove(pim, pim + 32 * 64, 8 * 64 * sizeof(float));
    ref_esbr_anal32(tin, anal + (size_t)u * 320, apos + 2 * u, qa);
    for (int s = 0; s < 32; s++) {
      memcpy(qre + 64 * (40 + s), qa + 128 * s, 32 * sizeof(float));
      memcpy(qim + 64 * (40 + s), qa + 128 * s + 64, 32 * sizeof(float));
    }
    e |= ref_esbr_hbe_apply_tbl(hbe_cfg + (size_t)u * XO_HBE_CFG_WORDS, hbe_state + (size_t)u * XO_HBE_ST_WORDS, qre + 40 * 64,
                                qim + 40 * 64, pre + 8 * 64, pim + 8 * 64, hbe_tbl);
    memcpy(lre, qre, sizeof(lre));
    memcpy(lim, qim, sizeof(lim));
    e |= ref_esbr_generate_hf(lre, lim, pre, pim, ore, oim, hf_par + (size_t)u * XO_EHF_PAR_WORDS, bw + 6 * u, patch + 8 * u);
    e |= ref_esbr_env_calc(ore, oim, ec_ipar + (size_t)u * XO_EEC_IPAR_WORDS, ec_fpar + (size_t)u * XO_EEC_FPAR_WORDS,
                           ec + (size_t)u * 640);
    const int32_t *r = rg + 4 * u;
    for (int s = 0; s < 32; s++) {
      const int xo = s < r[2] ? r[0] : r[1];
      for (int k = 0; k < 64; k++) {
        m[128 * s + k] = k < xo ? qre[64 * (2 + s) + k] : ore[64 * (2 + s) + k];
        m[128 * s + 64 + k] = k < xo ? qim[64 * (2 + s) + k] : oim[64 * (2 + s) + k];
      }
    }
    ref_esbr_synth64(m, synth + (size_t)u * 1280, spos + 2 * u, tout);
    int16_t *o = pcm + (size_t)(u / 2) * 4096 + (u & 1);
    for (int i = 0; i < 2048; i++) {
      float v = tout[i];
      if (v > 32767.0f) v = 32767.0f; else if (v < -32768.0f) v = -32768.0f;
      o[2 * i] = (int16_t)v;
    }
    err[u] = e;
  }
}
#endif

/* ixheaacd_qmf_hbe_data_reinit (hbe_trans.c:103-222) on flat frequency tables, for test-side construction of cfg[] */
int ref_esbr_hbe_reinit(const int16_t *tbl_lo, int num_lo, const int16_t *tbl_hi, int num_hi, int32_t *cfg) {
  static ia_esbr_hbe_txposer_struct t;
  static double pers[16384];
  WORD32 used = 0;
  WORD16 lo[64], hi[64], nsfb[2] = {(WORD16)num_lo, (WORD16)num_hi};
  WORD16 *tab[2] = {lo, hi};
  memset(pers, 0, sizeof(pers));
  ixheaacd_esbr_hbe_data_init(&t, 1024, 0, 2048, pers, &used);
  memcpy(lo, tbl_lo, 2 * (num_lo + 1));
  memcpy(hi, tbl_hi, 2 * (num_hi + 1));
  int ret = ixheaacd_qmf_hbe_data_reinit(&t, tab, nsfb, 0);
  hbe_cfg_pack(cfg, &t, 0);
  return ret;
}

#ifdef XAAC_REF_TAPS
/* ---- tap (linked into xaacdec_tap only) ---- */
int32_t g_hbe_cfg[XO_HBE_CFG_WORDS];
int g_hbe_called;
WORD32 __real_ixheaacd_qmf_hbe_apply(ia_esbr_hbe_txposer_struct *t, FLOAT32 a[][64], FLOAT32 b[][64], WORD32 num_columns,
                                     FLOAT32 c[][64], FLOAT32 d[][64], WORD32 pitch_in_bins, ia_sbr_header_data_struct *hd);
WORD32 __wrap_ixheaacd_qmf_hbe_apply(ia_esbr_hbe_txposer_struct *t, FLOAT32 qre[][64], FLOAT32 qim[][64], WORD32 num_columns,
                                     FLOAT32 pre[][64], FLOAT32 pim[][64], WORD32 pitch_in_bins,
                                     ia_sbr_header_data_struct *hd) {
  static FILE *fp = NULL;
  static int tried = 0, count = 0;
  if (!tried) {
    tried = 1;
    const char *p = getenv("XAAC_TAP_FILE"), *s = getenv("XAAC_TAP_STAGES");
    if (p && *p && s && strstr(s, "hbe")) {
      char name[1024];
      snprintf(name, sizeof(name), "%s.hbe", p);
      fp = fopen(name, "wb");
    }
  }
  const char *m = getenv("XAAC_TAP_MAX");
  const int rec = fp && count < (m ? atoi(m) : 1000000) && num_columns == 32 && t->no_bins == 32;
  /* the instance may be re-initialised inside the call (hbe_trans.c:240-248), so its sizes are only known afterwards:
   * snapshot the raw arrays first, cut the flat state_in once synth_size is known */
  static float raw_in[1088], raw_synth[1280], raw_anal[640], raw_qin[12][128], raw_qout[32][128];
  static float st_in[XO_HBE_ST_WORDS], st_out[XO_HBE_ST_WORDS];
  int32_t cfg[XO_HBE_CFG_WORDS], head[2];
  const int fft_null = t->ixheaacd_cmplx_anal_fft == NULL;
  if (rec) {
    memcpy(raw_in, t->ptr_input_buf, sizeof(raw_in));
    memcpy(raw_synth, t->synth_buf, sizeof(raw_synth));
    memcpy(raw_anal, t->analy_buf, sizeof(raw_anal));
    for (int r = 0; r < 12; r++) memcpy(raw_qin[r], t->qmf_in_buf[16 + r], 512);
    for (int r = 0; r < 32; r++) memcpy(raw_qout[r], t->qmf_out_buf[32 + r], 512);
  }
  WORD32 ret = __real_ixheaacd_qmf_hbe_apply(t, qre, qim, num_columns, pre, pim, pitch_in_bins, hd);
  if (getenv("XAAC_TAP_DEBUG"))
    fprintf(stderr, "hbe call: S %d k_start %d bands %d..%d stretch %d xo %d %d %d %d pitch %d fft_null %d ret %d\n", t->synth_size,
            t->k_start, t->start_band, t->end_band, t->max_stretch, t->x_over_qmf[0], t->x_over_qmf[1], t->x_over_qmf[2],
            t->x_over_qmf[3], pitch_in_bins, fft_null, ret);
  hbe_cfg_pack(g_hbe_cfg, t, pitch_in_bins); /* for the whole-stage tap (ref_taps_esbr.c) */
  g_hbe_called++;
  if (rec) {
    const int S = t->synth_size;
    int bad = S < 1 || S > 20;
    hbe_cfg_pack(cfg, t, pitch_in_bins);
    cfg[XO_HBE_REINIT] = fft_null;
    memset(st_in, 0, sizeof(st_in));
    if (!bad) {
      memcpy(st_in + XO_HBE_ST_TAIL, raw_in + 32 * S, S * 4);
      memcpy(st_in + XO_HBE_ST_SYNTH, raw_synth, 18 * S * 4);
      memcpy(st_in + XO_HBE_ST_ANAL, raw_anal, 18 * S * 4);
      memcpy(st_in + XO_HBE_ST_QIN, raw_qin, sizeof(raw_qin));
      memcpy(st_in + XO_HBE_ST_QOUT, raw_qout, 10 * 512);
      for (int r = 10; r < 32; r++)
        for (int c = 0; c < 128; c++)
          if (raw_qout[r][c] != 0.0f) bad = 1;
      bad |= hbe_state_pack(st_out, t);
    }
    head[0] = 0x31454248;
    head[1] = bad ? -101 : ret;
    fwrite(head, 4, 2, fp);
    fwrite(cfg, 4, XO_HBE_CFG_WORDS, fp);
    fwrite(st_in, 4, XO_HBE_ST_WORDS, fp);
    fwrite(qre, 4, 2048, fp);
    fwrite(qim, 4, 2048, fp);
    fwrite(pre, 4, 2048, fp);
    fwrite(pim, 4, 2048, fp);
    fwrite(st_out, 4, XO_HBE_ST_WORDS, fp);
    fflush(fp);
    count++;
  }
  return ret;
}
#endif
