/*
 * oracle/src/peaklim.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Restatement of the AAC-LC output stage of libxaac (SURVEY.md §8a-F "LC output"): ixheaacd_peak_limiter_process
 * (decoder/ixheaacd_peak_limiter.c:177-307, the WORD32 variant with PEAK_LIM_THR_FIX) followed by the round16 loop of
 * ixheaacd_dec_execute (decoder/ixheaacd_api.c:3676-3681).  The per-sample loop of the reference is split into the
 * phases the GPU kernel runs (scaled samples and channel maximum; the sliding-window maximum with the reference's own
 * index bookkeeping; raw gains; the gain smoothing recursion; delayed output), which is an identity transformation:
 * no phase feeds back into an earlier one.  Pinned against the compiled reference by tests/test_oracle_peaklim.py.
 * All float / double operations are single IEEE operations in the reference's order (gcc x86-64 SSE2, no FMA).
 */
#include <math.h>
#include <string.h>
#include "fixmath.h"
#include "xaac_oracle.h"

typedef union { float f; int32_t i; uint32_t u; } f32u;

/* st: XO_PL_* record (32-bit words); samples: interleaved WORD32 [frame_len][ch], in place; qshift_adj [ch];
 * pcm16 (optional): round16 of the result, interleaved.  Returns 0, or -1 for states the stage does not support. */
int xo_peak_limiter(int32_t *st, int32_t *samples, int frame_len, const int8_t *qshift_adj, int16_t *pcm16) {
  const int ch = st[XO_PL_NUM_CH], A = st[XO_PL_ATTACK];
  if (ch < 1 || ch > 2 || A < 1 || A > XO_PL_MAX_ATTACK || frame_len > 1024 || !st[XO_PL_LIMITER_ON]) return -1;
  float *max_buf = (float *)(st + XO_PL_MAX_BUF), *delayed = (float *)(st + XO_PL_DELAYED);
  f32u ac, rc, gm;
  ac.i = st[XO_PL_ATTACK_CONST]; rc.i = st[XO_PL_RELEASE_CONST]; gm.i = st[XO_PL_GAIN_MOD];
  double psg;
  memcpy(&psg, st + XO_PL_PSG, 8);
  int cir = st[XO_PL_CIR], max_idx = st[XO_PL_MAX_IDX], di = st[XO_PL_DELAY_IDX];
  static __thread float v[1024 * 2], t[1024], g[1024];
  /* phase 1: scaled samples and their channel maximum (peak_limiter.c:201-206, 261-265) */
  for (int i = 0; i < frame_len; i++) {
    float m = 0.0f;
    for (int j = 0; j < ch; j++) {
      const float gain_t = (float)(1 << qshift_adj[j]);
      const float x = (float)samples[i * ch + j] * gain_t;
      v[i * ch + j] = x;
      const float a = fabsf(x);
      m = m > a ? m : a; /* MAX(tmp, fabs(..)) */
    }
    t[i] = m;
  }
  /* phase 2: sliding maximum over the last A values with the reference's index bookkeeping (:207-222), phase 3: raw gain */
  const float thr = (float)2147483647; /* PEAK_LIM_THR_FIX converted for the comparison and the division */
  for (int i = 0; i < frame_len; i++) {
    max_buf[cir] = t[i];
    if (max_idx == cir) {
      max_idx = 0;
      for (int j = 1; j < A; j++)
        if (max_buf[j] > max_buf[max_idx]) max_idx = j;
    } else if (t[i] >= max_buf[max_idx]) {
      max_idx = cir;
    }
    if (++cir == A) cir = 0;
    const float mx = max_buf[max_idx];
    g[i] = mx > thr ? thr / mx : 1.0f;
  }
  /* phase 4: attack / release smoothing (:230-249) */
  float min_gain = 1.0f;
  for (int i = 0; i < frame_len; i++) {
    const float gain = g[i];
    if ((double)gain < psg) {
      const float c = (gain - 0.1f * (float)psg) * 1.11111111f;
      gm.f = gm.f > c ? c : gm.f;
    } else {
      gm.f = gain;
    }
    if ((double)gm.f < psg) {
      psg = (double)ac.f * (psg - (double)gm.f) + (double)gm.f;
      psg = psg > (double)gain ? psg : (double)gain;
    } else {
      psg = (double)rc.f * (psg - (double)gm.f) + (double)gm.f;
    }
    g[i] = (float)psg;
    if (g[i] < min_gain) min_gain = g[i];
  }
  /* phase 5: delayed output (:251-276) + round16 (api.c:3676-3681) */
  for (int i = 0; i < frame_len; i++) {
    for (int j = 0; j < ch; j++) {
      float x = delayed[di * ch + j];
      delayed[di * ch + j] = v[i * ch + j];
      x *= g[i];
      int64_t q = (int64_t)x;
      if (q > 2147483647LL) q = 2147483647LL;
      else if (q < -2147483647LL) q = -2147483647LL;
      samples[i * ch + j] = (int32_t)q;
      if (pcm16) pcm16[i * ch + j] = ox_round16((int32_t)q);
    }
    if (++di >= A) di = 0;
  }
  st[XO_PL_GAIN_MOD] = gm.i;
  memcpy(st + XO_PL_PSG, &psg, 8);
  f32u mg; mg.f = min_gain;
  st[XO_PL_MIN_GAIN] = mg.i;
  st[XO_PL_CIR] = cir; st[XO_PL_MAX_IDX] = max_idx; st[XO_PL_DELAY_IDX] = di;
  return 0;
}

void xo_peak_limiter_batch(int32_t *st, int32_t *samples, const int8_t *qshift_adj, int16_t *pcm16, int32_t *err, int ch, int n) {
  for (int u = 0; u < n; u++)
    err[u] = xo_peak_limiter(st + (size_t)u * XO_PL_WORDS, samples + (size_t)u * 1024 * ch, 1024, qshift_adj + (size_t)u * ch,
                             pcm16 ? pcm16 + (size_t)u * 1024 * ch : NULL);
}
