// sbr_glue_kernels.cu — the block-floating-point bookkeeping of the fixed-point HQ SBR stage for sm_100a (B200).
//
// Three small warp-per-unit kernels that sit between the heavy kernels of the stage and replace, bit-exactly, the
// orchestration code of ixheaacd_sbr_dec (decoder/ixheaacd_sbr_dec.c:662-1310, fixed branch, low_pow_flag = 0):
//   sbr_pre_kernel    :749-774   overlap slots -> matrix rows 0..5, ixheaacd_rescale_x_overlap
//                                (decoder/ixheaacd_sbrdec_lpfuncs.c:453-527)
//   sbr_scale_kernel  :1050-1127 headroom scans (ixheaacd_expsubbandsamples), the three ixheaacd_adjust_scale calls,
//                                scale-factor update, ixheaacd_clr_subsamples; builds the HF generator's argument record
//   sbr_post_kernel   :1205-1245, :1284-1308  previous-frame data, LPC state rows, overlap save, synthesis parameters
// All matrix accesses are row segments (lanes = consecutive bands), i.e. coalesced 128-byte requests.
#include <cstdint>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"
#include "sbr_common.cuh"

namespace xb {

constexpr int kGlueWarps = 8;

__global__ void __launch_bounds__(kGlueWarps * 32) sbr_pre_kernel(SbrStageArgs p) {
  const int lane = threadIdx.x & 31;
  const long long warps_total = (long long)gridDim.x * kGlueWarps;
  for (long long u = (long long)blockIdx.x * kGlueWarps + (threadIdx.x >> 5); u < p.n_units; u += warps_total) {
    i32 *m = p.matrix + u * kSbrMatWords;
    const i32 *ov = p.ov + u * 768;
    int16_t *sf = p.sf + u * 8, *misc = p.misc + u * 16;
    const int16_t *env = p.side + u * kSideWords + kSideEnv;
    {  // 3 KB copy: all six 16-byte requests of a lane in flight (the compiler cannot order m[] stores past ov[] loads itself)
      const int4 *src = reinterpret_cast<const int4 *>(ov);
      int4 *dst = reinterpret_cast<int4 *>(m);
      int4 v[6];
#pragma unroll
      for (int q = 0; q < 6; q++) v[q] = __ldg(src + lane + 32 * q);
#pragma unroll
      for (int q = 0; q < 6; q++) dst[lane + 32 * q] = v[q];
    }
    __syncwarp();
    if (p.side[u * kSideWords + kSideApply]) {
      const int old_lsb = misc[kMiscMaxQmfPrev], new_lsb = env[kEnvMaxQmfSubband];
      const int start_slot = env[kEnvTimeStep] * (misc[kMiscEndPosPrev] - env[kEnvNumTimeSlots]);
      const int syn_usb = misc[kMiscSynUsb];
      int ov_lb = sf[kSfOvLb], ov_hb = sf[kSfOvHb];
      __syncwarp();
      if (lane == 0) {
        misc[kMiscCodecUsb] = (int16_t)new_lsb;
        misc[kMiscSynLsb] = (int16_t)new_lsb;
      }
      if (new_lsb != old_lsb && old_lsb > 0) {
        int b0 = min(old_lsb, new_lsb), b1 = max(old_lsb, new_lsb);
        const int nz = new_lsb - old_lsb;
        if (nz > 0)
          for (int i = lane; i < (6 - start_slot) * nz; i += 32) {
            const int l = start_slot + i / nz, k = old_lsb + i % nz;
            m[128 * l + k] = 0;
            m[128 * l + 64 + k] = 0;
          }
        __syncwarp();
        int source_scale, target_scale, t_lsb, t_usb;
        if (new_lsb > old_lsb) { source_scale = ov_hb; target_scale = ov_lb; t_lsb = 0; t_usb = old_lsb; }
        else { source_scale = ov_lb; target_scale = ov_hb; t_lsb = old_lsb; t_usb = syn_usb; }
        const int reserve = warp_headroom(m, b0, b1, 0, start_slot, lane);
        warp_adjust_scale(m, b0, b1, 0, start_slot, reserve, lane);
        __syncwarp();
        source_scale += reserve;
        int delta = target_scale - source_scale;
        if (delta > 0) {
          delta = -delta;
          b0 = t_lsb;
          b1 = t_usb;
          if (lane == 0) sf[new_lsb > old_lsb ? kSfOvLb : kSfOvHb] = (int16_t)source_scale;
        }
        warp_adjust_scale(m, b0, b1, 0, start_slot, delta, lane);
      }
    }
    __syncwarp();
    if (lane == 0) p.usb[u] = misc[kMiscCodecUsb];
  }
}

__global__ void __launch_bounds__(kGlueWarps * 32) sbr_scale_kernel(SbrStageArgs p) {
  const int lane = threadIdx.x & 31;
  const long long warps_total = (long long)gridDim.x * kGlueWarps;
  for (long long u = (long long)blockIdx.x * kGlueWarps + (threadIdx.x >> 5); u < p.n_units; u += warps_total) {
    i32 *m = p.matrix + u * kSbrMatWords;
    i32 *lpc = p.lpc + u * 256;
    int16_t *sf = p.sf + u * 8;
    const int16_t *misc = p.misc + u * 16;
    const int16_t *side = p.side + u * kSideWords;
    const int usb = misc[kMiscCodecUsb];
    int reserve = warp_headroom(m, 0, usb, 6, 38, lane);
    int reserve_ov1 = warp_headroom(m, 0, usb, 0, 6, lane);
    const int reserve_ov2 = warp_headroom(lpc, 0, usb, 0, 2, lane);
    reserve_ov1 = min(reserve_ov1, reserve_ov2);
    const int lb0 = -8;  // set by the analysis stage (generic:635)
    const int ov_lb0 = sf[kSfOvLb];
    const int shift1 = lb0 + reserve, shift2 = ov_lb0 + reserve_ov1;
    const int min_shift = min(shift1, shift2);
    const int shift_over = shift2 - min_shift;
    reserve -= shift1 - min_shift;
    const int ov_shift = reserve_ov1 - shift_over;
    __syncwarp();
    warp_adjust_scale(m, 0, usb, 0, 6, ov_shift, lane);
    warp_adjust_scale(lpc, 0, usb, 0, 2, ov_shift, lane);
    {  // rows 6..37: shift bands < usb, clear bands 32..63 (ixheaacd_clr_subsamples, sbr_dec.c:1117-1127)
      const int sh = max(-31, min(31, reserve));
      for (int l = 6; l < 38; l++) {
        i32 *row = m + 128 * l;
        if (lane < usb && sh != 0) {
          const i32 a = row[lane], b = row[64 + lane];
          row[lane] = sh > 0 ? lsl(a, sh) : (a >> -sh);
          row[64 + lane] = sh > 0 ? lsl(b, sh) : (b >> -sh);
        }
        row[32 + lane] = 0;
        row[96 + lane] = 0;
      }
    }
    const int ov_lb = ov_lb0 + ov_shift, lb = lb0 + reserve;
    if (lane == 0) {
      sf[kSfStLb] = 0;
      sf[kSfOvLb] = (int16_t)ov_lb;
      sf[kSfLb] = (int16_t)lb;
      if (!side[kSideApply]) sf[kSfHb] = (int16_t)lb;  // sbr_dec.c:1215
    }
    // argument record of the HF generator (kernels.h kHf*): static transposer settings + derived scalars
    int16_t *hf = p.hf_prm + u * 80;
    const int16_t *env = side + kSideEnv, *hfs = side + kSideHf;
    for (int i = lane; i < 80; i += 32) {
      int v = hfs[i];
      if (i == kHfFactor) v = env[kEnvTimeStep];
      else if (i == kHfStartIdx) v = env[kEnvBorderVec];
      else if (i == kHfStopIdx) v = sat16(env[kEnvBorderVec + env[kEnvNumEnv]] - env[kEnvNumTimeSlots]);
      else if (i >= kHfInvfPrev && i < kHfInvfPrev + 10) v = misc[kMiscInvfPrev + (i - kHfInvfPrev)];
      else if (i == kHfOvLbScale) v = ov_lb;
      else if (i == kHfLbScale) v = lb;
      else if (i == kHfMaxQmfSubband) v = env[kEnvMaxQmfSubband];
      hf[i] = (int16_t)v;
    }
  }
}

__global__ void __launch_bounds__(kGlueWarps * 32) sbr_post_kernel(SbrStageArgs p) {
  const int lane = threadIdx.x & 31;
  const long long warps_total = (long long)gridDim.x * kGlueWarps;
  for (long long u = (long long)blockIdx.x * kGlueWarps + (threadIdx.x >> 5); u < p.n_units; u += warps_total) {
    const i32 *m = p.matrix + u * kSbrMatWords;
    int16_t *sf = p.sf + u * 8, *misc = p.misc + u * 16;
    const int16_t *side = p.side + u * kSideWords;
    const int16_t *env = side + kSideEnv, *hfs = side + kSideHf;
    int16_t *synp = p.synp + u * 8;
    if (p.err && p.err[u] != 0) {  // the reference returns before any of this (sbr_dec.c:1203)
      if (lane < 8) synp[lane] = 0;
      continue;
    }
    if (side[kSideApply]) {
      const int nif = hfs[kHfNumIfBands];
      if (lane < nif && lane < 10) misc[kMiscInvfPrev + lane] = hfs[kHfInvf + lane];
      if (lane == 0) {
        misc[kMiscMaxQmfPrev] = env[kEnvMaxQmfSubband];
        misc[kMiscEndPosPrev] = env[kEnvBorderVec + env[kEnvNumEnv]];
      }
    }
    const int usb = misc[kMiscCodecUsb];
    i32 *lpc = p.lpc + u * 256;
    // sbr_dec.c:1284-1290 copies 64 * op_delay = 384 words: slots 32..34 in the complex layout.  All loads of both copies
    // are issued before the first store (the compiler cannot move m[] loads past lpc[] / ov[] stores itself).
    i32 *ov = p.ov + u * 768;
    {
      i32 vl[4] = {0, 0, 0, 0};
      if (lane < usb) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          vl[2 * i] = m[128 * (30 + i) + lane];
          vl[2 * i + 1] = m[128 * (30 + i) + 64 + lane];
        }
      }
      const int4 *src = reinterpret_cast<const int4 *>(m + 32 * 128);
      int4 vo[3];
#pragma unroll
      for (int q = 0; q < 3; q++) vo[q] = src[lane + 32 * q];
      if (lane < usb) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          lpc[128 * i + lane] = vl[2 * i];
          lpc[128 * i + 64 + lane] = vl[2 * i + 1];
        }
      }
      int4 *dst = reinterpret_cast<int4 *>(ov);
#pragma unroll
      for (int q = 0; q < 3; q++) dst[lane + 32 * q] = vo[q];
    }
    __syncwarp();
    if (lane == 0) {
      synp[0] = sf[kSfOvLb];
      synp[1] = sf[kSfLb];
      synp[2] = sf[kSfHb];
      synp[3] = sf[kSfStSyn];
      synp[4] = misc[kMiscSynLsb];
      synp[5] = misc[kMiscSynUsb];
      synp[6] = 6;
      synp[7] = 0;
      sf[kSfOvLb] = sf[kSfLb];  // :1308 (save_lb_scale)
    }
  }
}

// Stage glue: WORD32 IMDCT output -> PCM16.
//   mode 0  ixheaacd_allocate_sbr_scr (decoder/ixheaacd_api.c:337-370): round16(shl32_sat(x, qshift_adj)) — SBR input
//   mode 1  ixheaacd_scale_adjust (decoder/ixheaacd_peak_limiter.c:324-333, wrapping x * (1 << qshift_adj)) + round16
//           (decoder/ixheaacd_api.c:3676-3681) — AAC-LC output with the peak limiter off
// Pure streaming: 16-byte loads, 8-byte stores, 6144 algorithmic bytes per unit.
__global__ void __launch_bounds__(256) pcm16_from_imdct_kernel(const int4 *in, const int8_t *qshift_adj, int2 *out,
                                                               long long n_vec, int mode) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
    const int q = qshift_adj[i >> 8];  // 256 vectors of 4 samples per unit
    const int4 v = __ldg(in + i);
    i32 a, b, c, d;
    if (mode) { a = lsl(v.x, q); b = lsl(v.y, q); c = lsl(v.z, q); d = lsl(v.w, q); }
    else { a = shl32_sat(v.x, q); b = shl32_sat(v.y, q); c = shl32_sat(v.z, q); d = shl32_sat(v.w, q); }
    int2 o;
    o.x = (round16(a) & 0xffff) | (i32)((u32)round16(b) << 16);
    o.y = (round16(c) & 0xffff) | (i32)((u32)round16(d) << 16);
    out[i] = o;
  }
}

cudaError_t launch_pcm16_from_imdct(const int32_t *in, const int8_t *qshift_adj, int16_t *out, long long n_units, int mode,
                                    int num_sms, cudaStream_t s) {
  const long long n_vec = n_units * 256;
  long long grid = (n_vec + 255) / 256;
  if (grid > (long long)num_sms * 16) grid = (long long)num_sms * 16;
  pcm16_from_imdct_kernel<<<(unsigned)grid, 256, 0, s>>>(reinterpret_cast<const int4 *>(in), qshift_adj,
                                                          reinterpret_cast<int2 *>(out), n_vec, mode);
  return cudaGetLastError();
}

static unsigned glue_grid(long long n_units, int num_sms) {
  long long need = (n_units + kGlueWarps - 1) / kGlueWarps;
  long long grid = (long long)num_sms * 8;
  if (grid > need) grid = need;
  return (unsigned)(grid < 1 ? 1 : grid);
}

cudaError_t launch_sbr_pre(const SbrStageArgs &a, int num_sms, cudaStream_t s) {
  sbr_pre_kernel<<<glue_grid(a.n_units, num_sms), kGlueWarps * 32, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_sbr_scale(const SbrStageArgs &a, int num_sms, cudaStream_t s) {
  sbr_scale_kernel<<<glue_grid(a.n_units, num_sms), kGlueWarps * 32, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_sbr_post(const SbrStageArgs &a, int num_sms, cudaStream_t s) {
  sbr_post_kernel<<<glue_grid(a.n_units, num_sms), kGlueWarps * 32, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace xb
