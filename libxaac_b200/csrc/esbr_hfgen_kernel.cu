// esbr_hfgen_kernel.cu — the float eSBR HF generator for sm_100a (B200).
//
// One warp owns one unit (one channel of one frame).  Replaces
//   ixheaacd_generate_hf              decoder/ixheaacd_sbrdec_lpfuncs.c:981-1359
//   ixheaacd_esbr_calc_co_variance    :781-830        ixheaacd_esbr_chirp_fac_calc   :832-849
// and ixheaacd_pre_processing :928-979 (+ ixheaacd_polyfit, ixheaacd_gausssolve :851-926)
// for the 2:1 system (38-slot covariance) without LD-MPS and error concealment.
// Every float operation is an explicit round-to-nearest intrinsic in the reference's evaluation order — the reference
// build has neither FMA contraction nor reassociation — so the float output is bit-identical, not just within 1 LSB.
//
// lane = QMF band everywhere: the covariance of band k is a serial sum over the time slots (the order matters), the 32
// lanes run 32 bands side by side and every global access is a coalesced 128-byte row segment.  The second-order
// prediction coefficients of the low band go through shared memory (a patch reads band k2 - stride); the patch walk is
// scalar and computed redundantly by all lanes.  No staging buffer: each source cell is read once for the covariance
// and once per patch that maps it, each destination cell is written once.
// Algorithmic HBM bytes per unit (typical 44.1 kHz tables: lsb 20, usb 46, slots 0..32): covariance 40 x 20 x 8 = 6.4 KB,
// patches 34 x 26 x 8 read + 32 x (64 - 20) x 8 written = 18.3 KB -> 24.7 KB.
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"

namespace xb {

constexpr int kEhWarps = 8;
constexpr int EHF_ROWS = 40;
constexpr size_t kEhUnit = (size_t)EHF_ROWS * 64;
#define EROW(b, i) ((b) + 64 * ((i) + 2))

struct EhWarpS {
  float a[4][64];  // alpha_real[k][0], alpha_imag[k][0], alpha_real[k][1], alpha_imag[k][1]
  i32 par[kEhfParWords];
  float bw[8];
  float gain[64];  // ixheaacd_pre_processing: gain per source band (1 without pre-processing); low-band level in dB on the way
};

struct Cov {
  float p01r, p01i, p02r, p02i, p11, p12r, p12i, p22, det;
};

// lpfuncs.c:781 — one band, slots 0..len-1 with two slots of history; the sliding registers hold the same values the
// reference re-reads from memory
XB_DEV void co_variance(Cov &c, const float *re, const float *im, int bd, int len) {
  c.p01r = c.p01i = c.p02r = c.p02i = c.p11 = c.p12r = c.p12i = c.p22 = 0.0f;
  float r2 = __ldg(EROW(re, -2) + bd), i2 = __ldg(EROW(im, -2) + bd);
  float r1 = __ldg(EROW(re, -1) + bd), i1 = __ldg(EROW(im, -1) + bd);
#pragma unroll 1
  for (int j0 = 0; j0 < len; j0 += 8) {
    float vr[8], vi[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int j = j0 + q < len ? j0 + q : len - 1;
      vr[q] = __ldg(EROW(re, j) + bd);
      vi[q] = __ldg(EROW(im, j) + bd);
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
      if (j0 + q < len) {
        const float r0 = vr[q], i0 = vi[q];
        c.p01r = __fadd_rn(c.p01r, __fadd_rn(__fmul_rn(r0, r1), __fmul_rn(i0, i1)));
        c.p01i = __fadd_rn(c.p01i, __fsub_rn(__fmul_rn(i0, r1), __fmul_rn(r0, i1)));
        c.p02r = __fadd_rn(c.p02r, __fadd_rn(__fmul_rn(r0, r2), __fmul_rn(i0, i2)));
        c.p02i = __fadd_rn(c.p02i, __fsub_rn(__fmul_rn(i0, r2), __fmul_rn(r0, i2)));
        c.p11 = __fadd_rn(c.p11, __fadd_rn(__fmul_rn(r1, r1), __fmul_rn(i1, i1)));
        c.p12r = __fadd_rn(c.p12r, __fadd_rn(__fmul_rn(r1, r2), __fmul_rn(i1, i2)));
        c.p12i = __fadd_rn(c.p12i, __fsub_rn(__fmul_rn(i1, r2), __fmul_rn(r1, i2)));
        c.p22 = __fadd_rn(c.p22, __fadd_rn(__fmul_rn(r2, r2), __fmul_rn(i2, i2)));
        r2 = r1; i2 = i1; r1 = r0; i1 = i0;
      }
    }
  }
  c.det = __fsub_rn(__fmul_rn(c.p11, c.p22),
                    __fmul_rn(__fadd_rn(__fmul_rn(c.p12r, c.p12r), __fmul_rn(c.p12i, c.p12i)), 0.999999f));
}

// lpfuncs.c:1056-1098 / 1264-1296
XB_DEV void solve_alpha(const Cov &c, float &ar0, float &ai0, float &ar1, float &ai1) {
  if (c.det == 0.0f) {
    ar1 = ai1 = 0.0f;
  } else {
    const float fac = __fdiv_rn(1.0f, c.det);
    ar1 = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(c.p01r, c.p12r), __fmul_rn(c.p01i, c.p12i)), __fmul_rn(c.p02r, c.p11)), fac);
    ai1 = __fmul_rn(__fsub_rn(__fadd_rn(__fmul_rn(c.p01i, c.p12r), __fmul_rn(c.p01r, c.p12i)), __fmul_rn(c.p02i, c.p11)), fac);
  }
  if (c.p11 == 0.0f) {
    ar0 = ai0 = 0.0f;
  } else {
    const float fac = __fdiv_rn(1.0f, c.p11);
    ar0 = __fmul_rn(-__fadd_rn(__fadd_rn(c.p01r, __fmul_rn(ar1, c.p12r)), __fmul_rn(ai1, c.p12i)), fac);
    ai0 = __fmul_rn(-__fsub_rn(__fadd_rn(c.p01i, __fmul_rn(ai1, c.p12r)), __fmul_rn(ar1, c.p12i)), fac);
  }
  const float m0 = __fadd_rn(__fmul_rn(ar0, ar0), __fmul_rn(ai0, ai0));
  const float m1 = __fadd_rn(__fmul_rn(ar1, ar1), __fmul_rn(ai1, ai1));
  if (m0 >= 16.0f || m1 >= 16.0f) ar0 = ai0 = ar1 = ai1 = 0.0f;
}

// ixheaacd_pre_processing + ixheaacd_polyfit + ixheaacd_gausssolve (decoder/ixheaacd_sbrdec_lpfuncs.c:851-979): the low band's
// level in dB per band (lane = band, serial over the slots in the reference's order), a third-order least-squares fit of it
// (one lane: the sums and the 4 x 4 elimination are serial by contract), gain[k] = 10^((mean - fit(k)) / 20).  log10 / pow run
// in double like the reference's libm calls and are rounded to float afterwards: CUDA's double-precision log10 / pow are within
// 1-2 ulp of the correctly rounded double, so the float differs from glibc's only when the double result lies within ~1e-16
// relative of a float rounding boundary (probability ~1e-8 per value; the tests allow 1 float ulp on such a gain).
__device__ __noinline__ void esbr_pre_processing(const float *sre, const float *sim, float *gain, int n, int start, int end,
                                                 int lane) {
  const unsigned full = 0xffffffffu;
  const bool have = n != 0 && end != start;
  for (int k = lane; k < 64; k += 32) {
    float le = 0.0f;
    if (have && k < n) {
      float temp = 0.0f;
#pragma unroll 4
      for (int i = start; i < end; i++) {
        const float r = __ldg(EROW(sre, i) + k), q = __ldg(EROW(sim, i) + k);
        temp = __fadd_rn(temp, __fadd_rn(__fmul_rn(r, r), __fmul_rn(q, q)));
      }
      temp = __fdiv_rn(temp, (float)(end - start));
      le = (float)(10.0 * log10((double)__fadd_rn(temp, 1.0f)));
    }
    gain[k] = le;  // low_env[k]
  }
  __syncwarp();
  float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f, mean = 0.0f;
  if (lane == 0) {
    if (have) {
      for (int k = 0; k < n; k++) mean = __fadd_rn(mean, gain[k]);
      mean = __fdiv_rn(mean, (float)n);
    }
    float a[4][4], b[4], y[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      b[i] = 0.0f;
#pragma unroll
      for (int j = 0; j < 4; j++) a[i][j] = 0.0f;
    }
    for (int k = 0; k < n; k++) {
      float v[7];
      v[0] = 1.0f;
#pragma unroll
      for (int i = 1; i <= 6; i++) v[i] = __fmul_rn((float)k, v[i - 1]);
      const float yk = gain[k];
#pragma unroll
      for (int i = 0; i <= 3; i++) {
        b[i] = __fadd_rn(b[i], __fmul_rn(v[3 - i], yk));
#pragma unroll
        for (int j = 0; j <= 3; j++) a[i][j] = __fadd_rn(a[i][j], v[6 - i - j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int imax = i;
#pragma unroll
      for (int k = i + 1; k < 4; k++) {
        float ak = 0.0f, am = 0.0f;  // a[k][i], a[imax][i] with static indices
#pragma unroll
        for (int r = 0; r < 4; r++) {
          if (r == k) ak = a[r][i];
          if (r == imax) am = a[r][i];
        }
        if (fabsf(ak) > fabsf(am)) imax = k;
      }
#pragma unroll
      for (int r = 0; r < 4; r++) {
        if (r == imax && r != i) {  // swap rows i and imax (columns >= i) and the right-hand sides
          const float t = b[r];
          b[r] = b[i];
          b[i] = t;
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (j >= i) {
              const float w = a[r][j];
              a[r][j] = a[i][j];
              a[i][j] = w;
            }
        }
      }
      const float v = a[i][i];
      b[i] = __fdiv_rn(b[i], v);
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (j >= i) a[i][j] = __fdiv_rn(a[i][j], v);
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (k > i) {
          const float w = a[k][i];
          b[k] = __fsub_rn(b[k], __fmul_rn(w, b[i]));
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (j > i) a[k][j] = __fsub_rn(a[k][j], __fmul_rn(w, a[i][j]));
        }
    }
#pragma unroll
    for (int i = 3; i >= 0; i--) {
      y[i] = b[i];
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (j > i) y[i] = __fsub_rn(y[i], __fmul_rn(a[i][j], y[j]));
    }
    p0 = y[0]; p1 = y[1]; p2 = y[2]; p3 = y[3];
  }
  p0 = __shfl_sync(full, p0, 0); p1 = __shfl_sync(full, p1, 0); p2 = __shfl_sync(full, p2, 0); p3 = __shfl_sync(full, p3, 0);
  mean = __shfl_sync(full, mean, 0);
  __syncwarp();
  for (int k = lane; k < 64; k += 32) {
    float g = 1.0f;
    if (k < n) {
      float x = (float)k, sl = p3;
      sl = __fadd_rn(sl, __fmul_rn(p2, x));
      x = __fmul_rn(x, x);
      sl = __fadd_rn(sl, __fmul_rn(p1, x));
      x = __fmul_rn(x, (float)k);
      sl = __fadd_rn(sl, __fmul_rn(p0, x));
      g = (float)pow(10.0, (double)__fdiv_rn(__fsub_rn(mean, sl), 20.0f));
    }
    gain[k] = g;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kEhWarps * 32) esbr_hfgen_kernel(EsbrHfgenArgs p) {
  __shared__ EhWarpS sm[kEhWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  EhWarpS &w = sm[warp];
  const long long warps_total = (long long)gridDim.x * kEhWarps;
  for (long long u = (long long)blockIdx.x * kEhWarps + warp; u < p.n_units; u += warps_total) {
    __syncwarp();
    for (int i = lane; i < kEhfParWords; i += 32) w.par[i] = p.par[u * kEhfParWords + i];
    __syncwarp();
    const i32 *par = w.par, *fm = w.par + kEhfFmaster, *invf_tbl = w.par + kEhfInvfTbl;
    const int num_mf = par[kEhfNumMf], num_if = par[kEhfNumIf], sb_start = par[kEhfSbStart];
    const int hbe = par[kEhfHbeFlag], patching = par[kEhfPatchingMode], fs = par[kEhfFs];
    int err = 0;
    if (par[kEhfUsf4] || num_mf < 1 || num_mf > 56 || num_if < 0 || num_if > 5 || fs <= 0) err = -2;
    const int lsb = err ? 0 : fm[0], usb = err ? 0 : fm[num_mf], xover = sb_start - lsb;
    const int start = par[kEhfBorderFirst] * 2, end = 32 + (par[kEhfBorderLast] - 16) * 2, cov_len = 38;
    if (start < 0 || end > EHF_ROWS - 2 || lsb < 0 || usb > 64 || lsb > usb) err = -2;
    for (int i = 0; i < 5; i++)
      if ((par[kEhfInvf + i] | par[kEhfInvfPrev + i]) & ~3) err = -2;
    if (err) {
      if (lane == 0 && p.err) p.err[u] = err;
      continue;
    }
    const float *sre = p.src_re + u * p.src_stride, *sim = p.src_im + u * p.src_stride;
    const float *pre = p.pv_re ? p.pv_re + u * kEhUnit : nullptr, *pim = p.pv_im ? p.pv_im + u * kEhUnit : nullptr;
    float *dre = p.dst_re + u * kEhUnit, *dim = p.dst_im + u * kEhUnit;
    float *bw_prev = p.bw_prev + 6 * u;
    if (p.shift_rows) {  // the stage's memmove of the 8 history rows (op_delay + SBR_HF_ADJ_OFFSET)
      float vr[16], vi[16];
#pragma unroll
      for (int q = 0; q < 16; q++) {
        vr[q] = dre[2048 + 32 * q + lane];
        vi[q] = dim[2048 + 32 * q + lane];
      }
#pragma unroll
      for (int q = 0; q < 16; q++) {
        dre[32 * q + lane] = vr[q];
        dim[32 * q + lane] = vi[q];
      }
      __syncwarp();
    }
    if (lane < 8) {  // lpfuncs.c:832
      float bw = 0.0f;
      if (lane < num_if) {
        const int mp = par[kEhfInvfPrev + lane] & 3, mc = par[kEhfInvf + lane] & 3;
        bw = mc == 3 ? 0.98f : mc == 2 ? 0.90f : mc == 1 ? (mp == 0 ? 0.60f : 0.75f) : (mp == 1 ? 0.60f : 0.00f);
        const float prev = bw_prev[lane];
        if (bw < prev)
          bw = __fadd_rn(__fmul_rn(0.75000f, bw), __fmul_rn(0.25000f, prev));
        else
          bw = __fadd_rn(__fmul_rn(0.90625f, bw), __fmul_rn(0.09375f, prev));
        if (bw < 0.015625f) bw = 0.0f;
      }
      w.bw[lane] = bw;
    }
    for (int k = usb + lane; k < 64; k += 32)
      for (int i = start; i < end; i++) EROW(dre, i)[k] = EROW(dim, i)[k] = 0.0f;
    for (int k = lane; k < 64; k += 32) w.a[0][k] = w.a[1][k] = w.a[2][k] = w.a[3][k] = 0.0f;
    __syncwarp();

    int patch = 0;
    const bool pre_proc = par[kEhfPreProc] != 0;
    if (pre_proc) esbr_pre_processing(sre, sim, w.gain, lsb, start, end, lane);  // lpfuncs.c:1052
    if (patching || !hbe) {
      int cov_count = lsb;
      if (par[kEhfMpsSbr]) cov_count = lsb < par[kEhfCovCount] ? lsb : par[kEhfCovCount];
      for (int k = lane; k < cov_count; k += 32) {
        if (k >= 1) {
          Cov c;
          co_variance(c, sre, sim, k, cov_len);
          solve_alpha(c, w.a[0][k], w.a[1][k], w.a[2][k], w.a[3][k]);
        }
      }
      __syncwarp();
      int goal_sb = __float2int_rz(__fadd_rn(__fdiv_rn(2.048e6f, (float)fs), 0.5f));
      if (goal_sb < usb) {
        int idx = 0;
        while (fm[idx] < goal_sb) idx++;
        goal_sb = fm[idx];
      } else {
        goal_sb = usb;
      }
      int src_start = xover + 1, sb = lsb + xover, flag_break = 0;
      i32 starts[7] = {0, 0, 0, 0, 0, 0, 0};
      while (sb < usb) {
        if (patch >= 6) {
          err = -1;
          break;
        }
#pragma unroll
        for (int q = 0; q < 6; q++)
          if (q == patch) starts[q] = sb;
        int nb = goal_sb - sb, stride;
        if (nb >= lsb - src_start) {
          stride = (sb - src_start) & ~1;
          nb = lsb - (sb - stride);
          const int goal = sb + nb;  // ixheaacd_find_closest_entry(direction 0), lpfuncs.c:263
          int v;
          if (goal <= fm[0]) {
            v = fm[0];
          } else if (goal >= fm[num_mf]) {
            v = fm[num_mf];
          } else {
            int idx = num_mf;
            while (fm[idx] > goal) idx--;
            v = fm[idx];
          }
          nb = v - sb;
        }
        stride = (nb + sb - lsb + 1) & ~1;
        src_start = 1;
        if (goal_sb - (sb + nb) < 3) goal_sb = usb;
        if (nb < 3 && patch > 0 && sb + nb == usb) {
          for (int k2 = sb + lane; k2 < sb + nb; k2 += 32)
            for (int i = start; i < end; i++) EROW(dre, i)[k2] = EROW(dim, i)[k2] = 0.0f;
          break;
        }
        if (nb < 0 && flag_break == 1) break;
        if (nb < 0) {
          flag_break = 1;
          continue;
        }
        flag_break = 0;
        bool bad2 = false, bad1 = false;
        for (int k2 = sb + lane; k2 < sb + nb; k2 += 32) {
          const int k = k2 - stride;
          if (k < 0 || k >= 64) {
            bad2 = true;
            continue;
          }
          int bwi = 0;
          while (bwi < 5 && k2 >= invf_tbl[bwi]) bwi++;
          if (bwi >= 5) {
            bad1 = true;
            continue;
          }
          float bw = w.bw[bwi];
          const float a0r = __fmul_rn(bw, w.a[0][k]), a0i = __fmul_rn(bw, w.a[1][k]);
          bw = __fmul_rn(bw, bw);
          const float a1r = __fmul_rn(bw, w.a[2][k]), a1i = __fmul_rn(bw, w.a[3][k]);
          const float gain = pre_proc ? w.gain[k] : 1.0f;
          if (bw > 0.0f) {
            float r2 = __ldg(EROW(sre, start - 2) + k), i2 = __ldg(EROW(sim, start - 2) + k);
            float r1 = __ldg(EROW(sre, start - 1) + k), i1 = __ldg(EROW(sim, start - 1) + k);
#pragma unroll 4
            for (int i = start; i < end; i++) {
              const float r0 = __ldg(EROW(sre, i) + k), i0 = __ldg(EROW(sim, i) + k);
              const float tr = __fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(a0r, r1), __fmul_rn(a0i, i1)), __fmul_rn(a1r, r2)),
                                         __fmul_rn(a1i, i2));
              const float ti = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0i, r1), __fmul_rn(a0r, i1)), __fmul_rn(a1i, r2)),
                                         __fmul_rn(a1r, i2));
              // dst = src * gain; dst += (...) * gain (lpfuncs.c:1226-1243); gain = 1 without pre-processing (exact products)
              EROW(dre, i)[k2] = __fadd_rn(__fmul_rn(r0, gain), __fmul_rn(tr, gain));
              EROW(dim, i)[k2] = __fadd_rn(__fmul_rn(i0, gain), __fmul_rn(ti, gain));
              r2 = r1; i2 = i1; r1 = r0; i1 = i0;
            }
          } else {
#pragma unroll 4
            for (int i = start; i < end; i++) {
              const float r0 = __ldg(EROW(sre, i) + k), i0 = __ldg(EROW(sim, i) + k);
              EROW(dre, i)[k2] = __fmul_rn(r0, gain);
              EROW(dim, i)[k2] = __fmul_rn(i0, gain);
            }
          }
        }
        if (__any_sync(0xffffffffu, bad2)) {
          err = -2;
          break;
        }
        if (__any_sync(0xffffffffu, bad1)) {
          err = -1;
          break;
        }
        sb += nb;
        patch++;
      }
      if (!err && lane == 0 && p.patch_out) {
#pragma unroll
        for (int q = 0; q < 7; q++) p.patch_out[8 * u + 1 + q] = starts[q];
      }
    }

    if (!err && pre && pim && hbe && !patching) {  // lpfuncs.c:1252-1337: the phase-vocoder (HBE) high band
      patch = 1;
      bool bad1 = false;
      for (int k2 = sb_start + lane; k2 < usb; k2 += 32) {
        Cov c;
        float ar0, ai0, ar1, ai1;
        co_variance(c, pre, pim, k2, cov_len);
        solve_alpha(c, ar0, ai0, ar1, ai1);
        int bwi = 0;  // the reference's index only ever advances; on a non-decreasing table that is the first fit from 0
        while (bwi < 5 && k2 >= invf_tbl[bwi]) bwi++;
        if (bwi >= 5) {
          bad1 = true;
          continue;
        }
        float bw = w.bw[bwi];
        const float a0r = __fmul_rn(bw, ar0), a0i = __fmul_rn(bw, ai0);
        bw = __fmul_rn(bw, bw);
        const float a1r = __fmul_rn(bw, ar1), a1i = __fmul_rn(bw, ai1);
        if (bw > 0.0f) {
          float r2 = __ldg(EROW(pre, start - 2) + k2), i2 = __ldg(EROW(pim, start - 2) + k2);
          float r1 = __ldg(EROW(pre, start - 1) + k2), i1 = __ldg(EROW(pim, start - 1) + k2);
#pragma unroll 4
          for (int i = start; i < end; i++) {
            const float r0 = __ldg(EROW(pre, i) + k2), i0 = __ldg(EROW(pim, i) + k2);
            const float tr = __fadd_rn(__fsub_rn(__fmul_rn(a0r, r1), __fmul_rn(a0i, i1)),
                                       __fsub_rn(__fmul_rn(a1r, r2), __fmul_rn(a1i, i2)));
            const float ti = __fadd_rn(__fadd_rn(__fmul_rn(a0i, r1), __fmul_rn(a0r, i1)),
                                       __fadd_rn(__fmul_rn(a1i, r2), __fmul_rn(a1r, i2)));
            EROW(dre, i)[k2] = __fadd_rn(r0, tr);
            EROW(dim, i)[k2] = __fadd_rn(i0, ti);
            r2 = r1; i2 = i1; r1 = r0; i1 = i0;
          }
        } else {
#pragma unroll 4
          for (int i = start; i < end; i++) {
            EROW(dre, i)[k2] = __ldg(EROW(pre, i) + k2);
            EROW(dim, i)[k2] = __ldg(EROW(pim, i) + k2);
          }
        }
      }
      if (__any_sync(0xffffffffu, bad1)) err = -1;
    }
    if (!err) {
      if (lane == 0 && p.patch_out) p.patch_out[8 * u] = patch;
      if (lane < num_if) bw_prev[lane] = w.bw[lane];
    }
    if (lane == 0 && p.err) p.err[u] = err;
  }
}

cudaError_t launch_esbr_hfgen(const EsbrHfgenArgs &args, int num_sms, cudaStream_t stream) {
  long long need = (args.n_units + kEhWarps - 1) / kEhWarps;
  static int occ = 0;  // persistent grid: exactly the CTAs that are resident, so "unit + warps_total" is the next one in flight
  if (!occ) {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, esbr_hfgen_kernel, kEhWarps * 32, 0);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
  }
  long long grid = (long long)num_sms * occ;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  esbr_hfgen_kernel<<<(unsigned)grid, kEhWarps * 32, 0, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace xb
