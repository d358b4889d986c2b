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
