/*
 * oracle/ref_shim_fps.c — TEST INFRASTRUCTURE ONLY.
 *
 * Flat-array entry points around the UNMODIFIED reference's float parametric stereo ixheaacd_esbr_apply_ps
 * (decoder/ixheaacd_ps_dec_flt.c:381) in the XAAC_FPS_* layouts of include/xaac_b200.h.  Compiled against the reference headers
 * where they lie; the struct <-> record conversions are the drop-in's own (libxaac_b200/dropin/ixheaacd_b200_pack_ps_flt.h),
 * so these entry points check both the kernels (GPU tests) and the host-side parameter preparation (CPU tests).
 *   par [n][REF_FPS_PAR_WORDS] int32: num_env, border_position[0..5], usb, iid_quant, freq_res_ipd, pad to 16,
 *                                     iid[5][20], icc[5][20], ipd[5][17], opd[5][17]
 *   hst [n][REF_FPS_HST_WORDS]: float h*_prev [8][20] (h11r h12r h21r h22r h11i h12i h21i h22i), int32 ipd_idx_map_1[17],
 *                               opd_idx_map_1[17], ipd_idx_map_2[17], opd_idx_map_2[17]
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
#include "ixheaacd_b200_pack_ps_flt.h"

#define REF_FPS_PAR_WORDS 386
#define REF_FPS_HST_WORDS 228

const void *ref_rom_fps_tables(int *bytes) {
  static float blob[XAAC_FPSROM_WORDS];
  b200_fps_pack_rom(blob, &ixheaacd_aac_dec_ps_tables, ixheaacd_aac_dec_ps_tables.rev_link_delay_ser);
  if (bytes) *bytes = (int)sizeof(blob);
  return blob;
}

static void hst_load(ia_ps_dec_struct *ps, const float *h) {
  float *prev[8] = {ps->h11_re_prev, ps->h12_re_prev, ps->h21_re_prev, ps->h22_re_prev,
                    ps->h11_im_prev, ps->h12_im_prev, ps->h21_im_prev, ps->h22_im_prev};
  const int32_t *m = (const int32_t *)(h + 160);
  for (int c = 0; c < 8; c++) memcpy(prev[c], h + 20 * c, 80);
  memcpy(ps->ipd_idx_map_1, m, 68); memcpy(ps->opd_idx_map_1, m + 17, 68);
  memcpy(ps->ipd_idx_map_2, m + 34, 68); memcpy(ps->opd_idx_map_2, m + 51, 68);
}
static void hst_store(const ia_ps_dec_struct *ps, float *h) {
  const float *prev[8] = {ps->h11_re_prev, ps->h12_re_prev, ps->h21_re_prev, ps->h22_re_prev,
                          ps->h11_im_prev, ps->h12_im_prev, ps->h21_im_prev, ps->h22_im_prev};
  int32_t *m = (int32_t *)(h + 160);
  for (int c = 0; c < 8; c++) memcpy(h + 20 * c, prev[c], 80);
  memcpy(m, ps->ipd_idx_map_1, 68); memcpy(m + 17, ps->opd_idx_map_1, 68);
  memcpy(m + 34, ps->ipd_idx_map_2, 68); memcpy(m + 51, ps->opd_idx_map_2, 68);
}

/* low_re / low_im [n][40][64]: left slot i = row 2 + i (rows 34..39 = the six look-ahead slots).  side_out / commit_out: what the
 * drop-in's b200_fps_side makes of the same parameters BEFORE the call (commit_out in the hst layout).  Returns 0, or -1 when
 * b200_fps_side refuses a unit (the reference is still run). */
int ref_fps_apply_batch(int64_t n, const float *low_re, const float *low_im, const int32_t *par, float *state, float *hst,
                        float *side_out, float *commit_out, float *left, float *right) {
  static ia_ps_dec_struct ps;
  static float lre[38][64], lim[38][64], rre[38][64], rim[38][64];
  float *plre[38], *plim[38], *prre[38], *prim[38];
  ia_ps_tables_struct *t = (ia_ps_tables_struct *)&ixheaacd_aac_dec_ps_tables;
  int rc = 0;
  for (int i = 0; i < 38; i++) { plre[i] = lre[i]; plim[i] = lim[i]; prre[i] = rre[i]; prim[i] = rim[i]; }
  for (int64_t u = 0; u < n; u++) {
    const int32_t *p = par + u * REF_FPS_PAR_WORDS;
    memset(&ps, 0, sizeof(ps));
    ixheaacd_create_ps_esbr_dec(&ps, t, 64, 32, 0);
    memcpy(ps.delay_sample_ser, t->rev_link_delay_ser, sizeof(ps.delay_sample_ser)); /* sbrdec_initfuncs.c:1054 */
    b200_fps_unpack_state(state + u * XAAC_FPS_ST_WORDS, &ps);
    hst_load(&ps, hst + u * REF_FPS_HST_WORDS);
    ps.num_env = (WORD16)p[0];
    for (int e = 0; e < 6; e++) ps.border_position[e] = (WORD16)p[1 + e];
    ps.iid_quant = p[8];
    ps.freq_res_ipd = p[9];
    for (int e = 0; e < 5; e++) {
      for (int b = 0; b < 20; b++) {
        ps.iid_par_table[e][b] = (WORD16)p[16 + 20 * e + b];
        ps.icc_par_table[e][b] = (WORD16)p[116 + 20 * e + b];
      }
      for (int b = 0; b < 17; b++) {
        ps.ipd_idx_map[e][b] = p[216 + 17 * e + b];
        ps.opd_idx_map[e][b] = p[301 + 17 * e + b];
      }
    }
    b200_fps_commit_rec cm;
    memset(&cm, 0, sizeof(cm));
    if (b200_fps_side(side_out + u * XAAC_FPS_SIDE_WORDS, &cm, &ps, t, p[7]) != 0) rc = -1;
    {
      float *c = commit_out + u * REF_FPS_HST_WORDS;
      int32_t *m = (int32_t *)(c + 160);
      memcpy(c, cm.h_last, 640);
      memcpy(m, cm.ipd1, 68); memcpy(m + 17, cm.opd1, 68); memcpy(m + 34, cm.ipd2, 68); memcpy(m + 51, cm.opd2, 68);
    }
    memcpy(lre, low_re + u * 2560 + 128, sizeof(lre));
    memcpy(lim, low_im + u * 2560 + 128, sizeof(lim));
    memset(rre, 0, sizeof(rre));
    memset(rim, 0, sizeof(rim));
    ixheaacd_esbr_apply_ps(&ps, plre, plim, prre, prim, p[7], t, 16);
    for (int i = 0; i < 32; i++) {
      memcpy(left + u * 4096 + 128 * i, lre[i], 256);
      memcpy(left + u * 4096 + 128 * i + 64, lim[i], 256);
      memcpy(right + u * 4096 + 128 * i, rre[i], 256);
      memcpy(right + u * 4096 + 128 * i + 64, rim[i], 256);
    }
    b200_fps_pack_state(state + u * XAAC_FPS_ST_WORDS, &ps);
    hst_store(&ps, hst + u * REF_FPS_HST_WORDS);
  }
  return rc;
}

#ifndef XAAC_REF_TAPS
/* The eSBR stage of a mono + PS element as the reference runs it with its default flags on a legacy HE-AACv2 stream (harmonic
 * transposer forced on, decoder/ixheaacd_sbrdecoder.c:400-403): analysis bank -> ixheaacd_qmf_hbe_apply -> ixheaacd_generate_hf ->
 * ixheaacd_sbr_env_calc -> regrouping + look-ahead slots -> ixheaacd_esbr_apply_ps -> synthesis bank for each output channel.
 * Built from the same per-stage entry points as ref_xheaac_hbe_chain_batch (oracle/ref_shim_hbe.c); units a..b-1, thread-safe.
 * q6 as there; ps_par / ps_hst in the layouts of ref_fps_apply_batch. */
void ref_esbr_anal32(const float *time_in, int32_t *states, int32_t *pos, float *qmf);
int ref_esbr_hbe_apply_tbl(const int32_t *cfg, float *state, const float *qmf_re, const float *qmf_im, float *pv_re, float *pv_im,
                           const int16_t *tbl);
int ref_esbr_generate_hf(const float *src_re, const float *src_im, const float *pv_re, const float *pv_im, float *dst_re,
                         float *dst_im, const int32_t *par, float *bw_prev, int32_t *patch_out);
int ref_esbr_env_calc(float *re, float *im, int32_t *ipar, const float *fpar, float *state);
void ref_esbr_synth64(const float *qmf, int32_t *fs, int32_t *pos, float *out);
#define FPS_Q6_WORDS (2 * 4608 + 4 * 2560)
void ref_heaacv2_esbr_chain_batch(const float *time_in, float *q6, int32_t *anal, int32_t *apos, int32_t *synth, int32_t *spos,
                                  float *bw, int32_t *patch, float *ec, float *hbe_state, const int32_t *hbe_cfg,
                                  const int16_t *hbe_tbl, const int32_t *hf_par, int32_t *ec_ipar, const float *ec_fpar,
                                  const int32_t *rg, const int32_t *ps_par, float *ps_state, float *ps_hst, int32_t *synth_r,
                                  int32_t *spos_r, float *out_l, float *out_r, int32_t *err, int a, int b) {
  static __thread ia_ps_dec_struct ps;
  static __thread float qa[32 * 128], m[32 * 128], lre[2560], lim[2560];
  static __thread float pl_re[38][64], pl_im[38][64], pr_re[38][64], pr_im[38][64];
  float *plre[38], *plim[38], *prre[38], *prim[38];
  ia_ps_tables_struct *t = (ia_ps_tables_struct *)&ixheaacd_aac_dec_ps_tables;
  for (int i = 0; i < 38; i++) { plre[i] = pl_re[i]; plim[i] = pl_im[i]; prre[i] = pr_re[i]; prim[i] = pr_im[i]; }
  for (int u = a; u < b; u++) {
    float *q = q6 + (size_t)u * FPS_Q6_WORDS;
    float *qre = q, *qim = q + 4608, *ore = q + 9216, *oim = ore + 2560, *pre = oim + 2560, *pim = pre + 2560;
    memmove(qre, qre + 32 * 64, 40 * 64 * sizeof(float));
    memmove(qim, qim + 32 * 64, 40 * 64 * sizeof(float));
    memmove(ore, ore + 32 * 64, 8 * 64 * sizeof(float));
    memmove(oim, oim + 32 * 64, 8 * 64 * sizeof(float));
    memmove(pre, pre + 32 * 64, 8 * 64 * sizeof(float));
    memmove(pim, pim + 32 * 64, 8 * 64 * sizeof(float));
    ref_esbr_anal32(time_in + (size_t)u * 1024, anal + (size_t)u * 320, apos + 2 * u, qa);
    for (int s = 0; s < 32; s++) {
      memcpy(qre + 64 * (40 + s), qa + 128 * s, 32 * sizeof(float));
      memcpy(qim + 64 * (40 + s), qa + 128 * s + 64, 32 * sizeof(float));
    }
    int e = ref_esbr_hbe_apply_tbl(hbe_cfg + (size_t)u * XAAC_HBE_CFG_WORDS, hbe_state + (size_t)u * XAAC_HBE_ST_WORDS,
                                   qre + 40 * 64, qim + 40 * 64, pre + 8 * 64, pim + 8 * 64, hbe_tbl);
    memcpy(lre, qre, sizeof(lre));
    memcpy(lim, qim, sizeof(lim));
    e |= ref_esbr_generate_hf(lre, lim, pre, pim, ore, oim, hf_par + (size_t)u * XAAC_EHF_PAR_WORDS, bw + 6 * u, patch + 8 * u);
    e |= ref_esbr_env_calc(ore, oim, ec_ipar + (size_t)u * XAAC_EEC_IPAR_WORDS, ec_fpar + (size_t)u * XAAC_EEC_FPAR_WORDS,
                           ec + (size_t)u * 640);
    const int32_t *r = rg + 4 * u;
    for (int s = 0; s < 32; s++) {
      const int xo = s < r[2] ? r[0] : r[1];
      for (int k = 0; k < 64; k++) {
        pl_re[s][k] = k < xo ? qre[64 * (2 + s) + k] : ore[64 * (2 + s) + k];
        pl_im[s][k] = k < xo ? qim[64 * (2 + s) + k] : oim[64 * (2 + s) + k];
      }
    }
    for (int s = 32; s < 38; s++)
      for (int k = 0; k < 5; k++) {
        pl_re[s][k] = qre[64 * (2 + s) + k];
        pl_im[s][k] = qim[64 * (2 + s) + k];
      }
    /* the PS instance is rebuilt from its flat state on every call, like the transposer's */
    const int32_t *p = ps_par + (size_t)u * REF_FPS_PAR_WORDS;
    ixheaacd_create_ps_esbr_dec(&ps, t, 64, 32, 0);
    memcpy(ps.delay_sample_ser, t->rev_link_delay_ser, sizeof(ps.delay_sample_ser));
    b200_fps_unpack_state(ps_state + (size_t)u * XAAC_FPS_ST_WORDS, &ps);
    hst_load(&ps, ps_hst + (size_t)u * REF_FPS_HST_WORDS);
    ps.num_env = (WORD16)p[0];
    for (int i = 0; i < 6; i++) ps.border_position[i] = (WORD16)p[1 + i];
    ps.iid_quant = p[8];
    ps.freq_res_ipd = p[9];
    for (int en = 0; en < 5; en++) {
      for (int bn = 0; bn < 20; bn++) {
        ps.iid_par_table[en][bn] = (WORD16)p[16 + 20 * en + bn];
        ps.icc_par_table[en][bn] = (WORD16)p[116 + 20 * en + bn];
      }
      for (int bn = 0; bn < 17; bn++) {
        ps.ipd_idx_map[en][bn] = p[216 + 17 * en + bn];
        ps.opd_idx_map[en][bn] = p[301 + 17 * en + bn];
      }
    }
    ixheaacd_esbr_apply_ps(&ps, plre, plim, prre, prim, ec_ipar[(size_t)u * XAAC_EEC_IPAR_WORDS + XAAC_EEC_SB_END], t, 16);
    b200_fps_pack_state(ps_state + (size_t)u * XAAC_FPS_ST_WORDS, &ps);
    hst_store(&ps, ps_hst + (size_t)u * REF_FPS_HST_WORDS);
    for (int s = 0; s < 32; s++) {
      memcpy(m + 128 * s, pl_re[s], 256);
      memcpy(m + 128 * s + 64, pl_im[s], 256);
    }
    ref_esbr_synth64(m, synth + (size_t)u * 1280, spos + 2 * u, out_l + (size_t)u * 2048);
    for (int s = 0; s < 32; s++) {
      memcpy(m + 128 * s, pr_re[s], 256);
      memcpy(m + 128 * s + 64, pr_im[s], 256);
    }
    ref_esbr_synth64(m, synth_r + (size_t)u * 1280, spos_r + 2 * u, out_r + (size_t)u * 2048);
    err[u] = e;
  }
}
#endif

/* b200_fps_side alone (the drop-in's host-side parameter preparation) on the shim's par / hst records; hst advances to what the
 * device call would commit.  Returns 0 or -1 (some unit refused). */
int ref_fps_side_batch(int64_t n, const int32_t *par, float *hst, float *side_out) {
  static ia_ps_dec_struct ps;
  ia_ps_tables_struct *t = (ia_ps_tables_struct *)&ixheaacd_aac_dec_ps_tables;
  int rc = 0;
  for (int64_t u = 0; u < n; u++) {
    const int32_t *p = par + u * REF_FPS_PAR_WORDS;
    memset(&ps, 0, sizeof(ps));
    ixheaacd_create_ps_esbr_dec(&ps, t, 64, 32, 0);
    hst_load(&ps, hst + u * REF_FPS_HST_WORDS);
    ps.num_env = (WORD16)p[0];
    for (int e = 0; e < 6; e++) ps.border_position[e] = (WORD16)p[1 + e];
    ps.iid_quant = p[8];
    ps.freq_res_ipd = p[9];
    for (int e = 0; e < 5; e++) {
      for (int b = 0; b < 20; b++) {
        ps.iid_par_table[e][b] = (WORD16)p[16 + 20 * e + b];
        ps.icc_par_table[e][b] = (WORD16)p[116 + 20 * e + b];
      }
      for (int b = 0; b < 17; b++) {
        ps.ipd_idx_map[e][b] = p[216 + 17 * e + b];
        ps.opd_idx_map[e][b] = p[301 + 17 * e + b];
      }
    }
    b200_fps_commit_rec cm;
    memset(&cm, 0, sizeof(cm));
    if (b200_fps_side(side_out + u * XAAC_FPS_SIDE_WORDS, &cm, &ps, t, p[7]) != 0) { rc = -1; continue; }
    b200_fps_commit(&ps, &cm, t);
    hst_store(&ps, hst + u * REF_FPS_HST_WORDS);
  }
  return rc;
}
