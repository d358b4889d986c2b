// sbr_glue_kernels.cu — the block-floating-point bookkeeping of the fixed-point HQ SBR stage for sm_100a (B200).
//
// Three small warp-per-unit kernels that sit between the heavy kernels of the stage and replace, bit-exactly, the
// orchestration code of ixheaacd_sbr_dec (decoder/ixheaacd_sbr_dec.c:662-1310, fixed branch, low_pow_flag = 0).  The
// per-unit bodies live in sbr_glue_units.cuh: the stage driver normally runs them inside the analysis / envelope
// kernels (sbr_front_hq_kernel, calc_sbrenvelope_hq_kernel) and launches these only with XAAC_B200_SBR_UNFUSED=1.
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
#include "sbr_glue_units.cuh"

namespace xb {

constexpr int kGlueWarps = 8;

__global__ void __launch_bounds__(kGlueWarps * 32) sbr_pre_kernel(SbrStageArgs p) {
  const int lane = threadIdx.x & 31;
  const long long warps_total = (long long)gridDim.x * kGlueWarps;
  for (long long u = (long long)blockIdx.x * kGlueWarps + (threadIdx.x >> 5); u < p.n_units; u += warps_total) {
    const int usb = sbr_pre_unit(p, u, lane);
    if (lane == 0) p.usb[u] = (int16_t)usb;
  }
}

__global__ void __launch_bounds__(kGlueWarps * 32) sbr_scale_kernel(SbrStageArgs p) {
  const int lane = threadIdx.x & 31;
  const long long warps_total = (long long)gridDim.x * kGlueWarps;
  for (long long u = (long long)blockIdx.x * kGlueWarps + (threadIdx.x >> 5); u < p.n_units; u += warps_total)
    sbr_scale_unit(p, u, lane, p.misc[u * 16 + kMiscCodecUsb], -1);
}

__global__ void __launch_bounds__(kGlueWarps * 32) sbr_post_kernel(SbrStageArgs p) {
  const int lane = threadIdx.x & 31;
  const long long warps_total = (long long)gridDim.x * kGlueWarps;
  for (long long u = (long long)blockIdx.x * kGlueWarps + (threadIdx.x >> 5); u < p.n_units; u += warps_total)
    sbr_post_unit(p, u, lane, p.err && p.err[u] != 0);
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
