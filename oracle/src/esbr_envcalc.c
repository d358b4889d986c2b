/* esbr_envcalc.c — TEST INFRASTRUCTURE ONLY: plain-C restatement of the float eSBR envelope adjuster
 *   ixheaacd_sbr_env_calc     decoder/ixheaacd_esbr_envcal.c:71-908, the ORIG_SBR branch (:611-860) and the epilogue (:864-907)
 * for the 2:1 system without reset / limiter-table rebuild (ixheaacd_createlimiterbands is control plane), PVC, LD-MPS,
 * inter-TES (gamma = 0 makes ixheaacd_apply_inter_tes a no-op) and error concealment.  float / double promotion of every
 * expression follows the reference (guard and the noise-floor ratio are doubles); pinned against the compiled function. */
#include <math.h>
#include <string.h>
#include "xaac_oracle.h"

#define ROW(b, i) ((b) + 64 * ((i) + 2))
#define EEC_EPS 1e-12f

static const float fir[5][5] = {{1.0f},
                                {0.33333333333333f, 0.66666666666666f},
                                {0.12500000000000f, 0.37500000000000f, 0.50000000000000f},
                                {0.05857864376269f, 0.20000000000000f, 0.34142135623731f, 0.40000000000000f},
                                {0.03183050093751f, 0.11516383427084f, 0.21816949906249f, 0.30150283239582f,
                                 0.33333333333333f}};
static const float lim_gains[4] = {0.70795f, 1.0f, 1.41254f, 1e10f};
static const float hphase[2][4] = {{1.0f, 0.0f, -1.0f, 0.0f}, {0.0f, 1.0f, 0.0f, -1.0f}};

int xo_esbr_env_calc(const float *rphase, float *re, float *im, int32_t *ipar, const float *fpar, float *state) {
  const int sbs = ipar[XO_EEC_SB_START], sbe = ipar[XO_EEC_SB_END], nsub = sbe - sbs;
  const int num_env = ipar[XO_EEC_NUM_ENV], trans_env = ipar[XO_EEC_TRANS_ENV];
  const int num_nf = ipar[XO_EEC_NUM_NF], int_mode = ipar[XO_EEC_INTERPOL_FREQ];
  const int lb = ipar[XO_EEC_LIMITER_BANDS], lg = ipar[XO_EEC_LIMITER_GAINS];
  const int smoothing_length = ipar[XO_EEC_SMOOTHING_MODE] ? 0 : 4;
  const int32_t *border = ipar + XO_EEC_BORDER, *freq_res = ipar + XO_EEC_FREQ_RES;
  const int32_t *nborder = ipar + XO_EEC_NOISE_BORDER, *tbl_noise = ipar + XO_EEC_TBL_NOISE;
  const int32_t *tbl[2] = {ipar + XO_EEC_TBL_LO, ipar + XO_EEC_TBL_HI};
  const int32_t num_sf[2] = {ipar[XO_EEC_NUM_SF_LO], ipar[XO_EEC_NUM_SF_HI]};
  const int32_t *gate_mode = ipar + XO_EEC_GATE_MODE, *lim = ipar + XO_EEC_LIM_TABLE + 13 * (lb & 3);
  int8_t *harm_prev = (int8_t *)(ipar + XO_EEC_HARM_PREV);
  const float *sfb_nrg = fpar + XO_EEC_SFB_NRG, *noise_floor = fpar + XO_EEC_NOISE_FLOOR;
  float(*e_gain)[64] = (float(*)[64])state, (*noise_buf)[64] = (float(*)[64])(state + 320);
  int harm_index = ipar[XO_EEC_HARM_INDEX], phase_index = ipar[XO_EEC_PHASE_INDEX], start_up = ipar[XO_EEC_START_UP];
  const double guard = 1e-17;
  int8_t harmonics[64];
  float nrg_tone[64], noise_level[64], nrg_est[64], nrg_ref[64], nrg_gain[64], tmpf[64];
  int kk = 0, next = -1, m = 0;

  if (ipar[XO_EEC_SBR_MODE] != 1 || ipar[XO_EEC_USF4]) return -2;
  /* envcal.c:169-190: the limiter tables of these frames are rebuilt first (ixheaacd_createlimiterbands, host side) */
  if ((ipar[XO_EEC_RESET] || ipar[XO_EEC_PATCHING_CHANGED]) && !ipar[XO_EEC_LIM_REBUILT]) return -2;
  if (ipar[XO_EEC_RESET]) { start_up = 1; phase_index = 0; }
  if (sbs < 0 || sbe > 64 || nsub < 0 || num_env < 1 || num_env > 8 || num_nf < 1 || num_nf > 5 || (lb & ~3) || (lg & ~3)) return -2;
  if (num_sf[0] < 0 || num_sf[0] > 28 || num_sf[1] < 0 || num_sf[1] > 56 || gate_mode[lb] < 0 || gate_mode[lb] > 12) return -2;
  if ((unsigned)harm_index > 3u || (unsigned)phase_index > 511u) return -2;
  for (int i = 0; i < num_env; i++)
    if (ipar[XO_EEC_INTER_TES + i] || border[i] < 0 || 2 * border[i + 1] > XO_EEC_NUM_ROWS_MAX) return -2;
  for (int c = 0; c <= gate_mode[lb]; c++)
    if (lim[c] < 0 || lim[c] > 64) return -2;

  memset(harmonics, 0, 64);
  for (int i = 0; i < num_sf[1]; i++) { /* envcal.c:612 */
    const int li = tbl[1][i], ui = tbl[1][i + 1];
    const int t = ((ui + li) - (sbs << 1)) >> 1;
    if (t >= 64 || t < 0) return -1;
    harmonics[t] = (int8_t)ipar[XO_EEC_ADD_HARM + i];
  }

  for (int i = 0; i < num_env; i++) {
    if (kk > 2) return (int)0x80000000;
    if (border[i] == nborder[kk]) kk++, next++;
    if (next < 0) return -2; /* the reference would index the noise floor at -num_nf */
    const int noise_absc = (i == trans_env || i == ipar[XO_EEC_SHORT_PREV]) ? 1 : 0;
    const int smooth_length = noise_absc ? 0 : smoothing_length;
    const float *sf = fir[smooth_length];
    const int res = freq_res[i] & 1, l0 = 2 * border[i], l1 = 2 * border[i + 1];
    int c = 0, o = 0;
    for (int j = 0; j < num_sf[res]; j++) { /* envcal.c:640 */
      const int li = tbl[res][j], ui = tbl[res][j + 1];
      int ui2 = tbl_noise[o + 1], flag = 0;
      float nrg;
      if (li < 0 || ui > 64 || ui < li || c + (ui - li) > 64) return -2;
      for (int k = li; k < ui; k++) {
        nrg = 0;
        if (l0 < l1) {
          for (int l = l0; l < l1; l++) nrg += (ROW(re, l)[k] * ROW(re, l)[k]) + (ROW(im, l)[k] * ROW(im, l)[k]);
          nrg = nrg / (l1 - l0);
        }
        if (harmonics[c] && (i >= trans_env || harm_prev[c + sbs])) flag = 1;
        nrg_est[c++] = nrg;
      }
      if (!int_mode && ui != li) {
        nrg = 0;
        for (int k = c - (ui - li); k < c; k++) nrg += nrg_est[k];
        nrg /= (ui - li);
      } else {
        nrg = 0;
      }
      c -= (ui - li);
      for (int k = 0; k < ui - li; k++) {
        double t;
        if (k + li >= ui2) o++;
        if (o >= 5) return (int)0x80000000;
        ui2 = tbl_noise[o + 1];
        const float nf = noise_floor[next * num_nf + o];
        nrg_ref[c] = sfb_nrg[m];
        if (!int_mode) nrg_est[c] = nrg;
        nrg_tone[c] = 0;
        t = nf / (1 + nf + guard);
        if (flag) {
          nrg_gain[c] = (float)sqrt(nrg_ref[c] * t / (nrg_est[c] + 1));
          if (harmonics[c] && (i >= trans_env || harm_prev[c + sbs])) nrg_tone[c] = (float)sqrt(nrg_ref[c] * t / fabs(nf + guard));
        } else if (noise_absc) {
          nrg_gain[c] = (float)sqrt(nrg_ref[c] / (nrg_est[c] + 1));
        } else {
          nrg_gain[c] = (float)sqrt(nrg_ref[c] * t / ((nrg_est[c] + 1) * fabs(nf + guard)));
        }
        noise_level[c] = (float)sqrt(nrg_ref[c] * t);
        c++;
      }
      m++;
    }

    for (int q = 0; q < gate_mode[lb]; q++) { /* envcal.c:726: limiter, then boost */
      float p_ref = 0, p_est = 0, p_adj = 0, avg_gain, g_max, boost;
      for (int k = lim[q]; k < lim[q + 1]; k++) {
        p_ref += nrg_ref[k];
        p_est += nrg_est[k];
      }
      avg_gain = (float)sqrt((p_ref + EEC_EPS) / (p_est + EEC_EPS));
      g_max = avg_gain * lim_gains[lg];
      if (g_max > 1.0e5f) g_max = 1.0e5f;
      for (int k = lim[q]; k < lim[q + 1]; k++)
        if (g_max <= nrg_gain[k]) {
          noise_level[k] = (float)(noise_level[k] * (g_max / (nrg_gain[k] + guard)));
          nrg_gain[k] = g_max;
        }
      for (int k = lim[q]; k < lim[q + 1]; k++) {
        p_adj += nrg_gain[k] * nrg_gain[k] * nrg_est[k];
        if (nrg_tone[k])
          p_adj += nrg_tone[k] * nrg_tone[k];
        else if (!noise_absc)
          p_adj += noise_level[k] * noise_level[k];
      }
      boost = (float)sqrt((p_ref + EEC_EPS) / (p_adj + EEC_EPS));
      if (boost > 1.584893192f) boost = 1.584893192f;
      for (int k = lim[q]; k < lim[q + 1]; k++) {
        nrg_gain[k] *= boost;
        noise_level[k] *= boost;
        nrg_tone[k] *= boost;
      }
    }

    if (start_up) {
      for (int n = 0; n < 4; n++) {
        memcpy(e_gain[n], nrg_gain, nsub * sizeof(float));
        memcpy(noise_buf[n], noise_level, nsub * sizeof(float));
      }
      start_up = 0;
    }

    for (int l = l0; l < l1; l++) { /* envcal.c:775 */
      float *pr = ROW(re, l) + sbs, *pi = ROW(im, l) + sbs;
      for (int k = 0; k < nsub; k++) {
        float sb_gain = 0, sb_noise = 0;
        int cc = 0;
        e_gain[4][k] = nrg_gain[k];
        noise_buf[4][k] = noise_level[k];
        for (int n = 4 - smooth_length; n <= 4; n++) {
          sb_gain += e_gain[n][k] * sf[cc];
          sb_noise += noise_buf[n][k] * sf[cc++];
        }
        phase_index = (phase_index + 1) & 511;
        if (nrg_tone[k] != 0 || noise_absc) sb_noise = 0;
        pr[k] = pr[k] * sb_gain + sb_noise * rphase[2 * phase_index];
        pi[k] = pi[k] * sb_gain + sb_noise * rphase[2 * phase_index + 1];
      }
      memcpy(tmpf, e_gain[0], sizeof(tmpf));
      memmove(e_gain[0], e_gain[1], 4 * sizeof(tmpf));
      memcpy(e_gain[4], tmpf, sizeof(tmpf));
      memcpy(tmpf, noise_buf[0], sizeof(tmpf));
      memmove(noise_buf[0], noise_buf[1], 4 * sizeof(tmpf));
      memcpy(noise_buf[4], tmpf, sizeof(tmpf));
    }
    /* ixheaacd_apply_inter_tes with gamma = 0: nothing.  Then the sinusoids (envcal.c:840) */
    for (int l = l0; l < l1; l++) {
      float *pr = ROW(re, l) + sbs, *pi = ROW(im, l) + sbs;
      int freq_inv = (sbs & 1) ? -1 : 1;
      for (int k = 0; k < nsub; k++) {
        pr[k] += nrg_tone[k] * hphase[0][harm_index];
        pi[k] += nrg_tone[k] * freq_inv * hphase[1][harm_index];
        freq_inv = -freq_inv;
      }
      harm_index = (harm_index + 1) & 3;
    }
  }

  memcpy(harm_prev + sbs, harmonics, 64 - sbs);
  ipar[XO_EEC_SHORT_PREV] = (trans_env == num_env) ? 0 : -1;
  if (ipar[XO_EEC_NUM_NOISE_ENV] < 1 || ipar[XO_EEC_NUM_NOISE_ENV] > 2) return (int)0x80000000;
  ipar[XO_EEC_HARM_INDEX] = harm_index;
  ipar[XO_EEC_PHASE_INDEX] = phase_index;
  ipar[XO_EEC_START_UP] = start_up;
  return 0;
}

void xo_esbr_env_calc_batch(const float *rphase, float *re, float *im, int32_t *ipar, const float *fpar, float *state,
                            int32_t *err, int n) {
  for (int u = 0; u < n; u++)
    err[u] = xo_esbr_env_calc(rphase, re + (size_t)u * 2560, im + (size_t)u * 2560, ipar + (size_t)u * XO_EEC_IPAR_WORDS,
                              fpar + (size_t)u * XO_EEC_FPAR_WORDS, state + (size_t)u * XO_EEC_STATE_WORDS);
}
