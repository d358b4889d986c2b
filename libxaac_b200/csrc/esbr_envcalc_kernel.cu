// esbr_envcalc_kernel.cu — the float eSBR envelope adjuster for sm_100a (B200).
//
// One warp owns one unit (one channel of one frame).  Replaces the ORIG_SBR branch of
//   ixheaacd_sbr_env_calc        decoder/ixheaacd_esbr_envcal.c:71-908  (:611-860 and the epilogue :864-907)
// for the 2:1 system without reset / limiter-table rebuild (ixheaacd_createlimiterbands is control plane: the host
// passes lim_table / gate_mode), PVC, LD-MPS, inter-TES (gamma = 0) and error concealment; other units get err = -2.
// The reference mixes float and double (the noise-floor ratio and the guard are doubles); every operation here is an
// explicit round-to-nearest intrinsic of the same type in the same order, so the adjusted QMF cells, the smoothing
// history and the indices are bit-identical.
//
// Phases per envelope (the warp walks the envelopes serially, as the smoothing history requires):
//   map      scalar walk over the scale-factor bands: index c -> absolute band, band of c, noise band of c
//   energy   lane = c: serial sum over the envelope's slots of |x|^2 (coalesced rows)
//   sfb      lane = scale-factor band: sinusoid flag, band-averaged energy (serial sum, order kept)
//   gain     lane = c: gain / noise level / sinusoid level in double, as the reference
//   limiter  lane = limiter band: serial power sums, limiting, boost
//   adjust   lane = band, slots serial: smoothing over the 5-deep history (registers), noise, sinusoids; in place
// Algorithmic HBM bytes per unit: slots x (sub_band_end - sub_band_start) x 8 read + the same written, + 5120 (history in /
// out) + 3008 (parameter records).
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"

namespace xb {

constexpr int kEcWarps = 8;
#define CROW(b, i) ((b) + 64 * ((i) + 2))

struct EcWarpS {
  i32 ipar[kEecIparWords];
  float nrg_est[64], nrg_ref[64], nrg_gain[64], noise_level[64], nrg_tone[64];
  float sfb_nrg[64];  // band-averaged energy per scale-factor band
  unsigned char kabs[64], sfb_of[64], o_of[64], hflag[64], harmonics[64], sfb_c0[64], sfb_flag[64];
};

__device__ __constant__ float kEcFir4[5] = {0.03183050093751f, 0.11516383427084f, 0.21816949906249f, 0.30150283239582f,
                                            0.33333333333333f};
__device__ __constant__ float kEcLimGains[4] = {0.70795f, 1.0f, 1.41254f, 1e10f};

__global__ void __launch_bounds__(kEcWarps * 32, 3) esbr_envcalc_kernel(EsbrEnvcalcArgs p) {
  __shared__ EcWarpS sm[kEcWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  EcWarpS &w = sm[warp];
  const long long warps_total = (long long)gridDim.x * kEcWarps;
  const double guard = 1e-17;
  for (long long u = (long long)blockIdx.x * kEcWarps + warp; u < p.n_units; u += warps_total) {
    __syncwarp();
    i32 *g_ipar = p.ipar + u * kEecIparWords;
    {
      i32 vi[kEecIparWords / 32];  // 9 requests in flight
#pragma unroll
      for (int q = 0; q < kEecIparWords / 32; q++) vi[q] = g_ipar[lane + 32 * q];
#pragma unroll
      for (int q = 0; q < kEecIparWords / 32; q++) w.ipar[lane + 32 * q] = vi[q];
    }
    __syncwarp();
    const i32 *ip = w.ipar;
    const int sbs = ip[kEecSbStart], sbe = ip[kEecSbEnd], nsub = sbe - sbs;
    const int num_env = ip[kEecNumEnv], trans_env = ip[kEecTransEnv], short_prev = ip[kEecShortPrev];
    const int num_nf = ip[kEecNumNf], int_mode = ip[kEecInterpolFreq];
    const int lb = ip[kEecLimiterBands], lg = ip[kEecLimiterGains];
    const int smoothing_length = ip[kEecSmoothingMode] ? 0 : 4;
    const i32 *border = ip + kEecBorder, *tbl_noise = ip + kEecTblNoise;
    const unsigned char *harm_prev = reinterpret_cast<const unsigned char *>(ip + kEecHarmPrev);
    int harm_index = ip[kEecHarmIndex], phase_index = ip[kEecPhaseIndex], start_up = ip[kEecStartUp];
    int err = 0;
    if (ip[kEecSbrMode] != 1 || ip[kEecUsf4]) err = -2;
    // envcal.c:169-190: on these frames the reference rebuilds its limiter tables from the patch table first; the host does
    // that (ixheaacd_createlimiterbands, control plane) and says so, otherwise the frame is refused
    if ((ip[kEecReset] || ip[kEecPatchingChanged]) && !ip[kEecLimRebuilt]) err = -2;
    if (ip[kEecReset]) { start_up = 1; phase_index = 0; }
    if (sbs < 0 || sbe > 64 || nsub < 0 || num_env < 1 || num_env > 8 || num_nf < 1 || num_nf > 5 || (lb & ~3) || (lg & ~3)) err = -2;
    const int num_sf_lo = ip[kEecNumSfLo], num_sf_hi = ip[kEecNumSfHi];
    const int gate = err ? 0 : ip[kEecGateMode + lb];
    if (num_sf_lo < 0 || num_sf_lo > 28 || num_sf_hi < 0 || num_sf_hi > 56 || gate < 0 || gate > 12) err = -2;
    if ((unsigned)harm_index > 3u || (unsigned)phase_index > 511u) err = -2;
    if (!err) {
      for (int i = 0; i < num_env; i++)
        if ((unsigned)ip[kEecInterTes + i] > 3u || (ip[kEecInterTes + i] && !p.low_re) || border[i] < 0 || 2 * border[i + 1] > 38)
          err = -2;
      const i32 *lim = ip + kEecLimTable + 13 * lb;
      for (int c = 0; c <= gate; c++)
        if (lim[c] < 0 || lim[c] > 64) err = -2;
    }
    if (err) {
      if (lane == 0 && p.err) p.err[u] = err;
      continue;
    }
    const i32 *lim = ip + kEecLimTable + 13 * lb;
    float *re = p.re + u * 2560, *im = p.im + u * 2560;
    const float *fpar = p.fpar + u * kEecFparWords;
    float *state = p.state + u * kEecStateWords;

    // harmonics (envcal.c:612): later scale-factor bands overwrite earlier ones, so the walk stays serial (all lanes alike)
    w.harmonics[lane] = 0;
    w.harmonics[lane + 32] = 0;
    __syncwarp();
    for (int i = 0; i < num_sf_hi; i++) {
      const int li = ip[kEecTblHi + i], ui = ip[kEecTblHi + i + 1];
      const int t = ((ui + li) - (sbs << 1)) >> 1;
      if (t >= 64 || t < 0) {
        err = -1;
        break;
      }
      if (lane == 0) w.harmonics[t] = (unsigned char)(ip[kEecAddHarm + i] != 0);
    }
    if (err) {
      if (lane == 0 && p.err) p.err[u] = err;
      continue;
    }
    // smoothing history of this lane's two bands
    float eg[2][5], nbuf[2][5];
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int n = 0; n < 5; n++) {
        eg[h][n] = state[64 * n + lane + 32 * h];
        nbuf[h][n] = state[320 + 64 * n + lane + 32 * h];
      }
    __syncwarp();

    int kk = 0, next = -1, m = 0, map_res = -1, C = 0;
    for (int i = 0; i < num_env && !err; i++) {
      if (kk > 2) {
        err = (int)0x80000000;
        break;
      }
      if (border[i] == ip[kEecNoiseBorder + kk]) kk++, next++;
      if (next < 0) {
        err = -2;
        break;
      }
      const int noise_absc = (i == trans_env || i == short_prev) ? 1 : 0;
      const int smooth_length = noise_absc ? 0 : smoothing_length;
      const int res = ip[kEecFreqRes + i] & 1, l0 = 2 * border[i], l1 = 2 * border[i + 1];
      const i32 *tbl = ip + (res ? kEecTblHi : kEecTblLo);
      const int num_sf = res ? num_sf_hi : num_sf_lo;

      // ---- map: c -> absolute band / scale-factor band / noise band (depends on the resolution only: kept across envelopes)
      if (res != map_res) {
        int o = 0;
        C = 0;
        map_res = res;
        for (int j = 0; j < num_sf; j++) {
          const int li = tbl[j], ui = tbl[j + 1];
          if (li < 0 || ui > 64 || ui < li || C + (ui - li) > 64) {
            err = -2;
            break;
          }
          int ui2 = tbl_noise[o + 1];
          if (lane == 0) w.sfb_c0[j] = (unsigned char)C;
          for (int k = 0; k < ui - li; k++) {
            if (k + li >= ui2) o++;
            if (o >= 5) {
              err = (int)0x80000000;
              break;
            }
            ui2 = tbl_noise[o + 1];
            if (lane == 0) {
              w.kabs[C + k] = (unsigned char)(li + k);
              w.sfb_of[C + k] = (unsigned char)j;
              w.o_of[C + k] = (unsigned char)o;
            }
          }
          if (err) break;
          C += ui - li;
        }
        if (lane == 0) w.sfb_c0[num_sf] = (unsigned char)C;
      }
      if (err) break;
      __syncwarp();

      // ---- energy: lane = c
      for (int c = lane; c < C; c += 32) {
        const int k = w.kabs[c];
        float nrg = 0.0f;
        if (l0 < l1) {
#pragma unroll 4
          for (int l = l0; l < l1; l++) {
            const float a = CROW(re, l)[k], b = CROW(im, l)[k];
            nrg = __fadd_rn(nrg, __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)));
          }
          nrg = __fdiv_rn(nrg, (float)(l1 - l0));
        }
        w.nrg_est[c] = nrg;
        w.hflag[c] = (unsigned char)(w.harmonics[c] && (i >= trans_env || (c + sbs < 64 && harm_prev[c + sbs])));
      }
      __syncwarp();
      // ---- per scale-factor band: sinusoid flag, averaged energy
      for (int j = lane; j < num_sf; j += 32) {
        const int c0 = w.sfb_c0[j], c1 = w.sfb_c0[j + 1];
        int flag = 0;
        float nrg = 0.0f;
        for (int c = c0; c < c1; c++) flag |= w.hflag[c];
        if (!int_mode && c1 != c0) {
          for (int c = c0; c < c1; c++) nrg = __fadd_rn(nrg, w.nrg_est[c]);
          nrg = __fdiv_rn(nrg, (float)(c1 - c0));
        }
        w.sfb_flag[j] = (unsigned char)flag;
        w.sfb_nrg[j] = nrg;
      }
      __syncwarp();
      // ---- gains: lane = c (envcal.c:690-722)
      for (int c = lane; c < C; c += 32) {
        const int j = w.sfb_of[c];
        const float nf = __ldg(fpar + kEecNoiseFloor + next * num_nf + w.o_of[c]);
        const float ref = __ldg(fpar + kEecSfbNrg + m + j);
        const float est = int_mode ? w.nrg_est[c] : w.sfb_nrg[j];
        const double t = __ddiv_rn((double)nf, __dadd_rn((double)__fadd_rn(1.0f, nf), guard));
        const double rt = __dmul_rn((double)ref, t);
        const float est1 = __fadd_rn(est, 1.0f);
        const double anf = fabs(__dadd_rn((double)nf, guard));
        float gain, tone = 0.0f;
        if (w.sfb_flag[j]) {
          gain = __double2float_rn(__dsqrt_rn(__ddiv_rn(rt, (double)est1)));
          if (w.hflag[c]) tone = __double2float_rn(__dsqrt_rn(__ddiv_rn(rt, anf)));
        } else if (noise_absc) {
          gain = __fsqrt_rn(__fdiv_rn(ref, est1));
        } else {
          gain = __double2float_rn(__dsqrt_rn(__ddiv_rn(rt, __dmul_rn((double)est1, anf))));
        }
        w.nrg_ref[c] = ref;
        w.nrg_est[c] = est;
        w.nrg_gain[c] = gain;
        w.nrg_tone[c] = tone;
        w.noise_level[c] = __double2float_rn(__dsqrt_rn(rt));
      }
      m += num_sf;
      __syncwarp();
      // ---- limiter + boost: lane = limiter band (envcal.c:726-760)
      if (lane < gate) {
        const int k0 = lim[lane], k1 = lim[lane + 1];
        float p_ref = 0.0f, p_est = 0.0f, p_adj = 0.0f;
        for (int k = k0; k < k1; k++) {
          p_ref = __fadd_rn(p_ref, w.nrg_ref[k]);
          p_est = __fadd_rn(p_est, w.nrg_est[k]);
        }
        const float avg_gain = __fsqrt_rn(__fdiv_rn(__fadd_rn(p_ref, 1e-12f), __fadd_rn(p_est, 1e-12f)));
        float g_max = __fmul_rn(avg_gain, kEcLimGains[lg]);
        if (g_max > 1.0e5f) g_max = 1.0e5f;
        for (int k = k0; k < k1; k++) {
          const float g = w.nrg_gain[k];
          if (g_max <= g) {
            w.noise_level[k] = __double2float_rn(
                __dmul_rn((double)w.noise_level[k], __ddiv_rn((double)g_max, __dadd_rn((double)g, guard))));
            w.nrg_gain[k] = g_max;
          }
        }
        for (int k = k0; k < k1; k++) {
          const float g = w.nrg_gain[k], tn = w.nrg_tone[k], nl = w.noise_level[k];
          p_adj = __fadd_rn(p_adj, __fmul_rn(__fmul_rn(g, g), w.nrg_est[k]));
          if (tn != 0.0f)
            p_adj = __fadd_rn(p_adj, __fmul_rn(tn, tn));
          else if (!noise_absc)
            p_adj = __fadd_rn(p_adj, __fmul_rn(nl, nl));
        }
        float boost = __fsqrt_rn(__fdiv_rn(__fadd_rn(p_ref, 1e-12f), __fadd_rn(p_adj, 1e-12f)));
        if (boost > 1.584893192f) boost = 1.584893192f;
        for (int k = k0; k < k1; k++) {
          w.nrg_gain[k] = __fmul_rn(w.nrg_gain[k], boost);
          w.noise_level[k] = __fmul_rn(w.noise_level[k], boost);
          w.nrg_tone[k] = __fmul_rn(w.nrg_tone[k], boost);
        }
      }
      __syncwarp();
      // ---- adjust: lane = band, slots serial (envcal.c:762-858)
      float gk[2], nk[2], tk[2], finv[2];
      bool act[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int k = lane + 32 * h;
        act[h] = k < nsub;
        gk[h] = act[h] ? w.nrg_gain[k] : 0.0f;
        nk[h] = act[h] ? w.noise_level[k] : 0.0f;
        tk[h] = act[h] ? w.nrg_tone[k] : 0.0f;
        finv[h] = (((sbs + k) & 1) ? -1.0f : 1.0f);
        if (start_up && act[h]) {
#pragma unroll
          for (int n = 0; n < 4; n++) {
            eg[h][n] = gk[h];
            nbuf[h][n] = nk[h];
          }
        }
      }
      start_up = 0;
      // inter-TES (envcal.c:824-836, 1021-1096) rescales the envelope's slots between the gain / noise pass and the sinusoids, so
      // an envelope that uses it (gamma > 0) adds its sinusoids in a pass of its own; otherwise both happen in one visit of a cell
      const int tes_mode = ip[kEecInterTes + i];
      const int harm_index0 = harm_index;
      for (int l = l0; l < l1; l++) {
        const float hp0 = tes_mode ? 0.0f : (harm_index == 0) ? 1.0f : (harm_index == 2) ? -1.0f : 0.0f;
        const float hp1 = tes_mode ? 0.0f : (harm_index == 1) ? 1.0f : (harm_index == 3) ? -1.0f : 0.0f;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          if (act[h]) {
            const int k = lane + 32 * h;
            eg[h][4] = gk[h];
            nbuf[h][4] = nk[h];
            float sb_gain = 0.0f, sb_noise = 0.0f;
            if (smooth_length == 0) {
              sb_gain = __fadd_rn(sb_gain, __fmul_rn(eg[h][4], 1.0f));
              sb_noise = __fadd_rn(sb_noise, __fmul_rn(nbuf[h][4], 1.0f));
            } else {
#pragma unroll
              for (int n = 0; n < 5; n++) {
                sb_gain = __fadd_rn(sb_gain, __fmul_rn(eg[h][n], kEcFir4[n]));
                sb_noise = __fadd_rn(sb_noise, __fmul_rn(nbuf[h][n], kEcFir4[n]));
              }
            }
            const int ph = (phase_index + k + 1) & 511;
            if (tk[h] != 0.0f || noise_absc) sb_noise = 0.0f;
            float *pr = CROW(re, l) + sbs + k, *pi = CROW(im, l) + sbs + k;
            float vr = __fadd_rn(__fmul_rn(*pr, sb_gain), __fmul_rn(sb_noise, __ldg(p.rphase + 2 * ph)));
            float vi = __fadd_rn(__fmul_rn(*pi, sb_gain), __fmul_rn(sb_noise, __ldg(p.rphase + 2 * ph + 1)));
            if (!tes_mode) {
              vr = __fadd_rn(vr, __fmul_rn(tk[h], hp0));
              vi = __fadd_rn(vi, __fmul_rn(__fmul_rn(tk[h], finv[h]), hp1));
            }
            *pr = vr;
            *pi = vi;
          }
          const float t0 = eg[h][0], t1 = nbuf[h][0];
#pragma unroll
          for (int n = 0; n < 4; n++) {
            eg[h][n] = eg[h][n + 1];
            nbuf[h][n] = nbuf[h][n + 1];
          }
          eg[h][4] = t0;
          nbuf[h][4] = t1;
        }
        phase_index = (phase_index + nsub) & 511;
        harm_index = (harm_index + 1) & 3;
      }
      __syncwarp();
      if (tes_mode) {
        const float gamma = tes_mode == 1 ? 1.0f : tes_mode == 2 ? 2.0f : 4.0f;  // ixheaac_q_gamma_table
        const int ns = l1 - l0;
        const float *lr0 = p.low_re + u * p.low_stride, *li0 = p.low_im + u * p.low_stride;
        // the low bands of the envelope's slots are copied next to the high bands first (envcal.c:1039-1044)
        for (int l = l0; l < l1; l++)
          for (int j = lane; j < sbs; j += 32) {
            CROW(re, l)[j] = CROW(lr0, l)[j];
            CROW(im, l)[j] = CROW(li0, l)[j];
          }
        __syncwarp();
        // per-slot powers, lane = slot (two rounds when the envelope has more than 32 slots), sums in band order
        float pl[2] = {0.f, 0.f}, ph[2] = {0.f, 0.f}, g[2] = {0.f, 0.f};
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int sI = lane + 32 * h;
          if (sI < ns) {
            const float *rr = CROW(re, l0 + sI), *ii = CROW(im, l0 + sI);
            float a = 0.0f, b = 0.0f;
            for (int j = 0; j < sbs; j++) {
              a = __fadd_rn(a, __fmul_rn(rr[j], rr[j]));
              a = __fadd_rn(a, __fmul_rn(ii[j], ii[j]));
            }
            for (int j = sbs; j < sbe; j++) {
              b = __fadd_rn(b, __fmul_rn(rr[j], rr[j]));
              b = __fadd_rn(b, __fmul_rn(ii[j], ii[j]));
            }
            pl[h] = a;
            ph[h] = b;
          }
        }
        float tot_lo = 0.0f, tot_hi = 0.0f;
        for (int sI = 0; sI < ns; sI++) {  // totals in slot order (every lane keeps a copy)
          tot_lo = __fadd_rn(tot_lo, __shfl_sync(0xffffffffu, sI < 32 ? pl[0] : pl[1], sI & 31));
          tot_hi = __fadd_rn(tot_hi, __shfl_sync(0xffffffffu, sI < 32 ? ph[0] : ph[1], sI & 31));
        }
        const float den = __fadd_rn(tot_lo, 1.0e-6f);
#pragma unroll
        for (int h = 0; h < 2; h++) {
          float q = (float)sqrt((double)__fdiv_rn(__fmul_rn(pl[h], (float)ns), den));
          q = __fadd_rn(1.0f, __fmul_rn(gamma, __fsub_rn(q, 1.0f)));
          if (q < 0.2f) q = 0.2f;
          g[h] = q;
          ph[h] = __fmul_rn(ph[h], __fmul_rn(q, q));
        }
        float tot_after = 1.0e-6f;
        for (int sI = 0; sI < ns; sI++)
          tot_after = __fadd_rn(tot_after, __shfl_sync(0xffffffffu, sI < 32 ? ph[0] : ph[1], sI & 31));
        const float gain_adj = (float)sqrt((double)__fdiv_rn(tot_hi, tot_after));
        for (int sI = 0; sI < ns; sI++) {
          const float gs = __fmul_rn(__shfl_sync(0xffffffffu, sI < 32 ? g[0] : g[1], sI & 31), gain_adj);
          float *rr = CROW(re, l0 + sI), *ii = CROW(im, l0 + sI);
          for (int j = sbs + lane; j < sbe; j += 32) {
            rr[j] = __fmul_rn(rr[j], gs);
            ii[j] = __fmul_rn(ii[j], gs);
          }
        }
        __syncwarp();
        // the sinusoids of the envelope (envcal.c:838-858)
        int hx = harm_index0;
        for (int l = l0; l < l1; l++) {
          const float hp0 = (hx == 0) ? 1.0f : (hx == 2) ? -1.0f : 0.0f;
          const float hp1 = (hx == 1) ? 1.0f : (hx == 3) ? -1.0f : 0.0f;
#pragma unroll
          for (int h = 0; h < 2; h++)
            if (act[h]) {
              const int k = lane + 32 * h;
              float *pr = CROW(re, l) + sbs + k, *pi = CROW(im, l) + sbs + k;
              *pr = __fadd_rn(*pr, __fmul_rn(tk[h], hp0));
              *pi = __fadd_rn(*pi, __fmul_rn(__fmul_rn(tk[h], finv[h]), hp1));
            }
          hx = (hx + 1) & 3;
        }
        __syncwarp();
      }
    }
    if (err) {
      if (lane == 0 && p.err) p.err[u] = err;
      continue;
    }
    // ---- epilogue: history, harmonic flags, indices
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int n = 0; n < 5; n++) {
        state[64 * n + lane + 32 * h] = eg[h][n];
        state[320 + 64 * n + lane + 32 * h] = nbuf[h][n];
      }
    {
      unsigned char *g_hp = reinterpret_cast<unsigned char *>(g_ipar + kEecHarmPrev);
      for (int k = lane; k < 64 - sbs; k += 32) g_hp[sbs + k] = w.harmonics[k];
    }
    if (lane == 0) {
      g_ipar[kEecShortPrev] = (trans_env == num_env) ? 0 : -1;
      const int nne = ip[kEecNumNoiseEnv];
      if (nne < 1 || nne > 2) {
        err = (int)0x80000000;
      } else {
        g_ipar[kEecHarmIndex] = harm_index;
        g_ipar[kEecPhaseIndex] = phase_index;
        g_ipar[kEecStartUp] = start_up;
      }
      if (p.err) p.err[u] = err;
    }
  }
}

cudaError_t launch_esbr_envcalc(const EsbrEnvcalcArgs &args, int num_sms, cudaStream_t stream) {
  static int occ = 0;
  if (!occ) {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, esbr_envcalc_kernel, kEcWarps * 32, 0);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
  }
  long long need = (args.n_units + kEcWarps - 1) / kEcWarps;
  long long grid = (long long)num_sms * occ;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  esbr_envcalc_kernel<<<(unsigned)grid, kEcWarps * 32, 0, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace xb
