// hfgen_kernel.cu — fixed-point complex ("HQ") SBR HF generator for sm_100a (B200).
//
// One warp owns one unit (one frame of one SBR channel); one lane owns one low band of the LPP transposer.
// Replaces, bit-exactly, the reference stage
//   ixheaacd_hf_generator                 decoder/ixheaacd_lpp_tran.c:956-1258
// and the leaves it calls:
//   ixheaacd_invfilt_level_emphasis       decoder/ixheaacd_sbrdec_lpfuncs.c:735-767
//   ixheaacd_covariance_matrix_calc_2_dec decoder/ixheaacd_lpp_tran.c:374-627     (selector: ixheaacd_covariance_matrix_calc_2)
//   ixheaacd_fix_div_dec                  decoder/ixheaacd_basic_funcs.c:130-152  (selector: ixheaacd_fix_div)
//   ixheaacd_filterstep3                  decoder/ixheaacd_lpp_tran.c:102-167
//
// The covariance sums wrap (plain adds under -fwrapv), so their closed forms (see oracle/src/hfgen.c) are evaluated
// in a single pass over the 40 rows with a 3-deep register window; a band column is read with one coalesced
// 128-byte request per row (lanes = consecutive bands).  Patches are built the same way: each lane streams its low
// band once more (L2 hits) through the 2-tap complex LPC filter and writes its high band, coalesced across lanes.
// Algorithmic HBM bytes per unit (HE-AACv2 tables, 16 low bands, 28 generated bands, 32 slots):
//   40 rows x 16 bands x 8 B read + 32 slots x 28 bands x 8 B written + 1 KB LPC rows + 160 B parameters ~= 13.5 KB.
#include <cstdint>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"

namespace xb {

constexpr int kHfWarps = 4;   // 38 KB of shared memory per block: five blocks = 20 warps per SM, four-warp blocks load the schedulers evenly (7-warp blocks: 1.26 -> 1.24 ms)
constexpr int kHfColWords = 38 * 2 * 32;
#ifndef HF_AHEAD
#define HF_AHEAD 8
#endif
constexpr int kHfAhead = HF_AHEAD;  // rows of the covariance pass in flight per lane (power of two); 4: 1.231 ms, 8: 1.167 ms, 16: 1.740 ms (128 registers)  // per warp: the low-band column set of one pass (38 rows x re, im x 32 bands)

// ops32.h:134 — second operand contributes only its high half (not commutative)
XB_DEV i32 hm(i32 a, i32 b) { return __mulhi(a, (i32)((u32)b & 0xffff0000u)); }
XB_DEV i32 abs_sat32(i32 a) { return a == (i32)0x80000000 ? 0x7fffffff : (a < 0 ? -a : a); }
XB_DEV i32 mult16_shl_sat(i32 a, i32 b) { return sat16((a * b) >> 15); }  // ops16.h:69

// basic_funcs.c:130-152
XB_DEV i32 fix_div(i32 op1, i32 op2) {
  i32 q = 0;
  i32 a = op1 >> 1, b = op2 >> 1;
  u32 num = (u32)(a < 0 ? -a : a), den = (u32)(b < 0 ? -b : b);
  if (num != 0) {
#pragma unroll 1
    for (int k = 15; k > 0; k--) {
      q <<= 1;
      num <<= 1;
      if (num >= den) {
        num -= den;
        q++;
      }
    }
  }
  return ((op1 ^ op2) < 0) ? -q : q;
}

__global__ void __launch_bounds__(kHfWarps * 32)
hf_generator_hq_kernel(HfGenArgs p) {
  __shared__ int16_t s_prm[kHfWarps][80];
  extern __shared__ __align__(16) i32 s_col[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // the covariance pass leaves every row of its band here; the patch passes (one per patch that maps the band) read it back
  // instead of going to L2 again
  i32 *colr = s_col + warp * kHfColWords + lane, *coli = colr + 38 * 32;
  const unsigned full = 0xffffffffu;
  const int warps_total = gridDim.x * kHfWarps;
  // ISO/IEC 14496-3 newBw table, Q31 (sbrdec_lpfuncs.c:85-89): rows = previous mode, columns = current mode
  const i32 nbw0 = 0x00000000, nbw06 = 0x4ccccccd, nbw075 = 0x60000000, nbw09 = 0x73333333, nbw098 = 0x7d70a3d7;

  for (long long u = (long long)blockIdx.x * kHfWarps + warp; u < p.n_units; u += warps_total) {
    if (p.gate && p.gate[u * p.gate_stride] == 0) continue;
    int16_t *prm = s_prm[warp];
    {
      const i32 *src = reinterpret_cast<const i32 *>(p.params + u * 80);
      for (int i = lane; i < 40; i += 32) reinterpret_cast<i32 *>(prm)[i] = __ldg(src + i);
    }
    __syncwarp();
    const int num_patches = prm[0], start_patch = prm[1], stop_patch = prm[2], num_columns = prm[3];
    const int factor = prm[50], num_if_bands = prm[51], max_qmf_subband = prm[76];
    const int L = (num_columns + 6 == 36) ? 36 : 38;
    const int start_idx = prm[52] * factor, stop_idx = num_columns + prm[53] * factor;
    i32 *mat = p.matrix + u * (38 * 128);
    const i32 *lpc = p.lpc + u * 256;

    // ---- inverse-filter level emphasis: lane i < num_if_bands owns bw_array[i] ----
    i32 my_bw = 0;
    if (lane < num_if_bands && lane < 6) {
      const int pm = prm[64 + lane], cm = prm[54 + lane];
      i32 nb = (cm == 3) ? nbw098 : (cm == 2) ? nbw09 : (cm == 1) ? ((pm == 0) ? nbw06 : nbw075)
                                                                  : ((pm == 1) ? nbw06 : nbw0);
      const i32 prev = p.bw_prev[u * 6 + lane];
      const i32 w1 = nb < prev ? 0x6000 : 0x7400, w2 = nb < prev ? 0x2000 : 0x0c00;
      i32 acc = wadd(lsl(mul32x16(nb, w1), 1), lsl(mul32x16(prev, w2), 1));
      if (acc < 0x02000000) acc = 0;
      if (acc >= 0x7f800000) acc = 0x7f800000;
      my_bw = acc;
      p.bw_prev[u * 6 + lane] = acc;
    }
    // ---- clear the bands above the last patch in the generated slots (lpp_tran.c:995-1007) ----
    {
      const int16_t *lp = prm + 14 + 6 * (num_patches - 1);
      const int actual_stop = (int16_t)(lp[3] + lp[5]);
      const int nb = 64 - actual_stop;
      if (nb > 0) {
        const int total = (stop_idx - start_idx) * nb;
        for (int i = lane; i < total; i += 32) {
          int row = start_idx + i / nb, b = actual_stop + i % nb;
          mat[128 * row + b] = 0;
          mat[128 * row + 64 + b] = 0;
        }
      }
    }
    if (lane == 0) {
      int cs = prm[74] < prm[75] ? prm[74] : prm[75];
      p.hb_scale[u * p.hb_stride] = (int16_t)(cs - 2);  // lpp_tran.c:1257, LPC_SCALE_FACTOR = 2
    }

    for (int lb0 = start_patch; lb0 < stop_patch; lb0 += 32) {
      const int lb = lb0 + lane;
      const bool active = lb < stop_patch;
      // ---- covariance of band lb over rows n = -2 .. L-1 (closed forms, see oracle/src/hfgen.c) ----
      i32 p01 = 0, p12 = 0, p01i = 0, p12i = 0, p11 = 0, p22 = 0, p02 = 0, p02i = 0;
      if (active) {
        i32 r2 = lpc[lb] >> 3, i2 = lpc[64 + lb] >> 3;               // n = -2
        i32 r1 = lpc[128 + lb] >> 3, i1 = lpc[128 + 64 + lb] >> 3;   // n = -1
        {  // contributions that only involve n = -2 / -1
          i32 C2 = wadd(hm(r2, r2), hm(i2, i2)), C1 = wadd(hm(r1, r1), hm(i1, i1));
          p22 = C2;
          p22 = wadd(p22, C1);
          p11 = C1;
          i32 A = wadd(hm(r1, r2), hm(i1, i2)), B = wsub(hm(i1, r2), hm(r1, i2));  // m = -1
          p12 = A;
          p12i = B;
        }
        // rows are prefetched kHfAhead ahead (register ring): each iteration's two loads are a coalesced 128-byte request per
        // component, and the arithmetic of row m overlaps the latency of rows m+1..m+4
        i32 pr[kHfAhead], pi[kHfAhead];
#pragma unroll
        for (int q = 0; q < kHfAhead; q++) { pr[q] = mat[128 * q + lb]; pi[q] = mat[128 * q + 64 + lb]; }
#pragma unroll kHfAhead
        for (int m = 0; m < L; m++) {
          constexpr int K = kHfAhead - 1;
          colr[32 * m] = pr[m & K];
          coli[32 * m] = pi[m & K];
          const i32 r0 = pr[m & K] >> 3, i0 = pi[m & K] >> 3;
          if (m + kHfAhead < L) { pr[m & K] = mat[128 * (m + kHfAhead) + lb]; pi[m & K] = mat[128 * (m + kHfAhead) + 64 + lb]; }
          i32 A = wadd(hm(r0, r1), hm(i0, i1)), B = wsub(hm(i0, r1), hm(r0, i1));
          p01 = wadd(p01, A);
          p01i = wadd(p01i, B);
          if (m <= L - 2) {
            p12 = wadd(p12, A);
            p12i = wadd(p12i, B);
            i32 C = wadd(hm(r0, r0), hm(i0, i0));
            p11 = wadd(p11, C);
            if (m <= L - 3) p22 = wadd(p22, C);
          }
          p02 = wadd(p02, wadd(hm(r0, r2), hm(i0, i2)));
          p02i = wadd(p02i, wsub(hm(i0, r2), hm(r0, i2)));
          r2 = r1; i2 = i1; r1 = r0; i1 = i0;
        }
      }
      // ---- LPC coefficients (lpp_tran.c:1050-1198) ----
      i32 ar0 = 0, ar1 = 0, ai0 = 0, ai1 = 0;
      if (active) {
        bool reset = false;
        i32 mx = abs_nrm(p01) | abs_nrm(p02) | abs_nrm(p12) | p11 | p22 | abs_nrm(p01i) | abs_nrm(p02i) | abs_nrm(p12i);
        int q = (mx == 0) ? 31 : (__clz(mx) - 1);  // pnorm32; mx >= 0 (all squared terms are non-negative)
        if (q < 0) q = 0;
        i32 c11 = lsl(p11, q), c22 = lsl(p22, q), c01 = lsl(p01, q), c02 = lsl(p02, q), c12 = lsl(p12, q);
        i32 c01i = lsl(p01i, q), c02i = lsl(p02i, q), c12i = lsl(p12i, q);
        i32 m2 = add_sat(mul32(c12, c12), mul32(c12i, c12i));
        i32 d = lsl(sub_sat(mul32(c11, c22), m2), 1);
        if (d != 0) {
          int nd = norm32(d);
          i32 inv = sext16(fix_div(0x40000000, lsl(d, nd)));
          i32 mod_d = abs_sat32(d);
          i32 tr = sub_sat(sub_sat(mul32(c01, c12), mul32(c01i, c12i)), mul32(c02, c11)) >> 1;
          i32 ti = sub_sat(add_sat(mul32(c01i, c12), mul32(c01, c12i)), mul32(c02i, c11)) >> 1;
          if (abs_sat32(tr) >= mod_d) reset = true;
          else ar1 = sext16(lsl(mul32x16(tr, inv), nd + 1) >> 15);
          if (abs_sat32(ti) >= mod_d) reset = true;
          else ai1 = sext16(lsl(mul32x16(ti, inv), nd + 1) >> 15);
        }
        if (c11 != 0) {
          int n11 = norm32(c11);
          i32 inv = sext16(fix_div(0x40000000, lsl(c11, n11)));
          i32 tr = add_sat(wadd(c01 >> 3, mul32x16(c12, ar1)), mul32x16(c12i, ai1));
          i32 ti = sub_sat(wadd(c01i >> 3, mul32x16(c12, ai1)), mul32x16(c12i, ar1));
          tr = lsl(tr, 1);
          ti = lsl(ti, 1);
          if (abs_sat32(tr) >= c11) reset = true;
          else ar0 = sext16(lsl(mul32x16(sub_sat(0, tr), inv), n11 + 1) >> 15);
          if (abs_sat32(ti) >= c11) reset = true;
          else ai0 = sext16(lsl(mul32x16(sub_sat(0, ti), inv), n11 + 1) >> 15);
        }
        if (add_sat(ar0 * ar0, ai0 * ai0) >= 0x40000000) reset = true;
        if (add_sat(ar1 * ar1, ai1 * ai1) >= 0x40000000) reset = true;
        if (reset) ar0 = ar1 = ai0 = ai1 = 0;
      }
      // ---- patches (lpp_tran.c:1200-1250); bw_index depends only on the high band (monotone scan) ----
      for (int pt = 0; pt < num_patches; pt++) {
        const int16_t *pp = prm + 14 + 6 * pt;
        const int hb = lb + pp[4];
        const bool go = active && lb >= pp[0] && lb < pp[1] && hb >= max_qmf_subband;
        int idx = 0;
        while (idx < 5 && hb >= prm[4 + idx]) idx++;
        const i32 bw32 = __shfl_sync(full, my_bw, idx);
        if (!go) continue;
        i32 bw = sext16(bw32 >> 16);
        const i32 a0r = mult16_shl_sat(bw, ar0), a0i = mult16_shl_sat(bw, ai0);
        bw = mult16_shl_sat(bw, bw);
        const i32 a1r = mult16_shl_sat(bw, ar1), a1i = mult16_shl_sat(bw, ai1);
        // rows relative to the reference scratch: row 0,1 = LPC states, row t+2 = matrix row t
        auto srcr = [&](int t) { return t < L ? colr[32 * t] : mat[128 * t + lb]; };  // matrix row t of band lb
        auto srci = [&](int t) { return t < L ? coli[32 * t] : mat[128 * t + 64 + lb]; };
        auto rowr = [&](int t) { return t < 2 ? lpc[128 * t + lb] : srcr(t - 2); };
        auto rowi = [&](int t) { return t < 2 ? lpc[128 * t + 64 + lb] : srci(t - 2); };
        i32 p2r = rowr(start_idx), p2i = rowi(start_idx), p1r = rowr(start_idx + 1), p1i = rowi(start_idx + 1);
#pragma unroll 2
        for (int t = start_idx; t < stop_idx; t++) {
          const i32 cr = srcr(t), ci = srci(t);  // scratch row t + 2
          i32 outr, outi;
          if (bw > 0) {
            i32 acc = wsub(wadd(wsub(mul32x16(p1r, a0r), mul32x16(p1i, a0i)), mul32x16(p2r, a1r)), mul32x16(p2i, a1i));
            outr = wadd(cr >> 2, lsl(acc, 1));
            acc = wadd(add_sat(add_sat(mul32x16(p1r, a0i), mul32x16(p1i, a0r)), mul32x16(p2r, a1i)), mul32x16(p2i, a1r));
            outi = wadd(ci >> 2, lsl(acc, 1));
          } else {
            outr = cr >> 2;
            outi = ci >> 2;
          }
          mat[128 * t + hb] = outr;
          mat[128 * t + 64 + hb] = outi;
          p2r = p1r; p2i = p1i; p1r = cr; p1i = ci;
        }
      }
    }
    __syncwarp();
  }
}

cudaError_t launch_hf_generator_hq(const HfGenArgs &args, int num_sms, cudaStream_t stream) {
  long long need = (args.n_units + kHfWarps - 1) / kHfWarps;
  long long grid = (long long)num_sms * 5;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  const size_t smem = (size_t)kHfWarps * kHfColWords * 4;
  static xb::PerDeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(hf_generator_hq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.done();
  }
  hf_generator_hq_kernel<<<(unsigned)grid, kHfWarps * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace xb
