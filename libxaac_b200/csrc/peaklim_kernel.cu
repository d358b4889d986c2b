// peaklim_kernel.cu — the AAC-LC output stage for sm_100a (B200): peak limiter + round16.
//
// One warp owns one stream (all channels of one 1024-sample frame; the limiter couples the channels through the common
// gain).  Replaces, bit-exactly,
//   ixheaacd_peak_limiter_process      decoder/ixheaacd_peak_limiter.c:177-307   (WORD32 variant, PEAK_LIM_THR_FIX)
//   the round16 loop of ixheaacd_dec_execute   decoder/ixheaacd_api.c:3676-3681
// for ia_peak_limiter_struct states produced by ixheaacd_peak_limiter_init (:45-75) with 1 or 2 channels.
//
// The reference's per-sample loop is split into phases that do not feed back into each other: (1) scaled samples and
// their channel maximum, lanes = samples; (2) the sliding maximum over the attack window with the reference's own
// (max_idx, cir_buf_pnt) bookkeeping — a uniform serial walk, the occasional window rescan is a warp argmax; (3) raw
// gains, lanes = samples; (4) the attack / release recursion — serial by nature, and skipped when the limiter is at
// rest and no sample of the frame exceeds the threshold (every real frame that does not clip); (5) delayed, limited
// output, lanes = samples.  All float / double arithmetic uses the round-to-nearest intrinsics in the reference's
// operation order (the reference build is x86-64 SSE2 without FMA), so results are bit-identical.
// Phase (4) is one dependent chain per STREAM.  Run inside this kernel it keeps a whole warp busy with one lane's worth of
// work (25 k of the kernel's 29 k warp instructions per stream), so streams that need it are DEFERRED when the caller
// provides scratch memory: this kernel stores their raw gains and queues them, peak_limiter_smooth_kernel runs the
// recursion with lane = stream (32 streams per warp, gains transposed through shared memory), and
// peak_limiter_finish_kernel emits their output (phase 5).  Streams at rest finish here as before.
// Algorithmic HBM bytes per stream (stereo): 8192 (WORD32 in) + 4096 (PCM16 out) + 2 x ~3 KB state.
#include <cstdint>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"

namespace xb {

constexpr int kPlWarps = 4;   // 41 KB of shared memory per block: five blocks = 20 warps per SM, evenly spread over the schedulers
constexpr int kPlSmoothWarps = 4;
constexpr int kPlFinishWarps = 8;

struct PlWarpS {
  float T[1024];   // channel maximum per sample
  float G[1024];   // window maximum, then raw gain, then smoothed gain per sample
  float mb[kPlMaxAttack];  // max_buf at frame start
};

// phase 5 (:251-276) + round16 (api.c:3676-3681): delayed input x smoothed gain G[i], clamp, lanes = samples; then the
// delay line takes the last A samples of this frame
XB_DEV void pl_emit_frame(const PeakLimArgs &p, long long u, i32 *st, const float *G, int lane, int ch, int A, int di0,
                          float gt0, float gt1) {
  const i32 *in = p.samples + u * 1024 * ch;
  float *dlg = reinterpret_cast<float *>(st + kPlDelayed);
  i32 *out32 = p.out32 ? p.out32 + u * 1024 * ch : nullptr;
  int16_t *pcm = p.pcm16 ? p.pcm16 + u * 1024 * ch : nullptr;
  auto lim = [](float x, float g) {
    const long long q = (long long)__fmul_rn(x, g);
    return (i32)(q > 2147483647LL ? 2147483647LL : (q < -2147483647LL ? -2147483647LL : q));
  };
  auto emit = [&](int i, int idx) {
    const float g = G[i];
    if (ch == 2) {  // both channels of a sample as one 8-byte request each way
      float x0, x1;
      if (i < A) {
        const float2 d = *reinterpret_cast<const float2 *>(dlg + 2 * idx);
        x0 = d.x, x1 = d.y;
      } else {
        const int2 v = *reinterpret_cast<const int2 *>(in + 2 * (i - A));
        x0 = __fmul_rn(__int2float_rn(v.x), gt0), x1 = __fmul_rn(__int2float_rn(v.y), gt1);
      }
      const i32 q0 = lim(x0, g), q1 = lim(x1, g);
      if (out32) *reinterpret_cast<int2 *>(out32 + 2 * i) = make_int2(q0, q1);
      if (pcm) *reinterpret_cast<i32 *>(pcm + 2 * i) = (round16(q0) & 0xffff) | (i32)((u32)round16(q1) << 16);
    } else {
      const float x = i < A ? dlg[idx] : __fmul_rn(__int2float_rn(in[i - A]), gt0);
      const i32 q = lim(x, g);
      if (out32) out32[i] = q;
      if (pcm) pcm[i] = (int16_t)round16(q);
    }
  };
  // samples 0 .. A-1 come out of the delay line, the others out of this frame; several requests in flight per lane
#pragma unroll 4
  for (int i = lane; i < A; i += 32) emit(i, (di0 + i) % A);
#pragma unroll 4
  for (int i = A + lane; i < 1024; i += 32) emit(i, 0);
  __syncwarp();  // every read of the old delay line is done before it is rewritten
#pragma unroll 1
  for (int i = 1024 - A + lane; i < 1024; i += 32) {
    const int idx = (di0 + i) % A;
    for (int j = 0; j < ch; j++) dlg[idx * ch + j] = __fmul_rn(__int2float_rn(in[i * ch + j]), j ? gt1 : gt0);
  }
}

// one step of the attack / release recursion (:230-249); returns the smoothed gain.  Branch-free: the two arms of the
// reference's second if differ only in the smoothing constant (and the attack arm's final max), so lanes that run
// different streams do not diverge.
XB_DEV float pl_smooth_step(float gain, float &gm, double &psg, const double ac, const double rc) {
  const double gd = (double)gain;
  const float c = __fmul_rn(__fsub_rn(gain, __fmul_rn(0.1f, (float)psg)), 1.11111111f);
  gm = gd < psg ? (gm > c ? c : gm) : gain;
  const double gmd = (double)gm;
  const bool attack = gmd < psg;
  double np = __dadd_rn(__dmul_rn(attack ? ac : rc, __dsub_rn(psg, gmd)), gmd);
  if (attack) np = np > gd ? np : gd;
  psg = np;
  return (float)np;
}

__global__ void __launch_bounds__(kPlWarps * 32) peak_limiter_kernel(PeakLimArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PlWarpS *ws = reinterpret_cast<PlWarpS *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;
  PlWarpS &w = ws[warp];
  const long long warps_total = (long long)gridDim.x * kPlWarps;
  const float thr = 2147483648.0f;  // (float)PEAK_LIM_THR_FIX

  for (long long u = (long long)blockIdx.x * kPlWarps + warp; u < p.n_units; u += warps_total) {
    __syncwarp();
    i32 *st = p.state + u * kPlWords;
    const int ch = st[kPlNumCh], A = st[kPlAttack];
    if (ch != p.ch || A < 1 || A > kPlMaxAttack || !st[kPlLimiterOn]) {
      if (lane == 0 && p.err) p.err[u] = (i32)0x80000000;
      continue;
    }
    const i32 *in = p.samples + u * 1024 * ch;
    float *mbg = reinterpret_cast<float *>(st + kPlMaxBuf);
    const float ac = __int_as_float(st[kPlAttackConst]), rc = __int_as_float(st[kPlReleaseConst]);
    float gm = __int_as_float(st[kPlGainMod]);
    double psg = __hiloint2double(st[kPlPsg + 1], st[kPlPsg]);
    int cir = st[kPlCir], max_idx = st[kPlMaxIdx];
    const int cir0 = cir, di0 = st[kPlDelayIdx];
    const float gt0 = (float)(1 << p.qshift_adj[u * ch]), gt1 = ch > 1 ? (float)(1 << p.qshift_adj[u * ch + 1]) : 0.f;
    for (int j = lane; j < A; j += 32) w.mb[j] = mbg[j];
    // ---- phase 1 (peak_limiter.c:201-206): t[i] = max_j |samples[i][j] * gain_t[j]| ----
#pragma unroll 4
    for (int i = lane; i < 1024; i += 32) {
      float m;
      if (ch == 2) {
        const int2 s = *reinterpret_cast<const int2 *>(in + 2 * i);
        m = fmaxf(fabsf(__fmul_rn(__int2float_rn(s.x), gt0)), fabsf(__fmul_rn(__int2float_rn(s.y), gt1)));
      } else {
        m = fabsf(__fmul_rn(__int2float_rn(in[i]), gt0));
      }
      w.T[i] = m;
    }
    __syncwarp();
    // ---- phase 2 (:207-222): window maximum with the reference's (max_idx, cir_buf_pnt) bookkeeping, 32 samples at a
    // time (lane = sample).  Inside a chunk the reference's serial walk has two kinds of events: (a) a sample >= the
    // running maximum moves max_idx to its own buffer position — found for all lanes at once with a prefix maximum;
    // (b) the write pointer reaches max_idx, i.e. the maximum leaves the window — the window is rescanned (warp argmax,
    // first maximum in buffer order).  After an (a) event max_idx is a position written in this chunk, which the pointer
    // cannot reach again before the chunk ends (attack window >= 32 samples), so (b) can only precede the first (a). ----
    float cur_max = w.mb[max_idx];
    if (A < 32) {  // shorter windows than a chunk (sample rates below 6.4 kHz): not produced by any supported stream
      if (lane == 0 && p.err) p.err[u] = (i32)0x80000000;
      continue;
    }
#pragma unroll 1
    for (int i0 = 0; i0 < 1024; i0 += 32) {
      const float tv = w.T[i0 + lane];
      int pos = cir + lane;
      if (pos >= A) pos -= A;
      int lane_start = 0;
      float mres = 0.f;
#pragma unroll 1
      while (true) {
        int lb = max_idx - cir;
        if (lb < 0) lb += A;
        const float tt = lane >= lane_start ? tv : -1.0f;
        float incl = tt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const float o = __shfl_up_sync(full, incl, d);
          if (lane >= d) incl = fmaxf(incl, o);
        }
        float excl = __shfl_up_sync(full, incl, 1);
        if (lane == 0) excl = -1.0f;
        excl = fmaxf(excl, cur_max);
        const unsigned aev = __ballot_sync(full, lane >= lane_start && tv >= excl);
        const int first_a = aev ? __ffs(aev) - 1 : 32;
        if (lb >= lane_start && lb < 32 && lb <= first_a) {
          if (lane >= lane_start && lane < lb) mres = cur_max;
          // the maximum was just overwritten at sample i0 + lb: rescan the window, first maximum in buffer order
          const int i = i0 + lb;
          int cirb = cir + lb;
          if (cirb >= A) cirb -= A;
          float bv = -1.0f;
          int bj = 0x7fffffff;
#pragma unroll 1
          for (int j = lane; j < A; j += 32) {
            int age = cirb - j;
            if (age < 0) age += A;
            const int ip = i - age;
            const float v = ip >= 0 ? w.T[ip] : w.mb[j];
            if (v > bv) { bv = v; bj = j; }
          }
          const int vm = __reduce_max_sync(full, __float_as_int(bv));  // values are >= 0: integer order = float order
          bj = (__float_as_int(bv) == vm) ? bj : 0x7fffffff;
          max_idx = __reduce_min_sync(full, bj);
          cur_max = __int_as_float(vm);
          if (lane == lb) mres = cur_max;
          lane_start = lb + 1;
          if (lane_start >= 32) break;
        } else {
          if (lane >= lane_start) mres = fmaxf(incl, cur_max);
          if (aev) {
            const int last_a = 31 - __clz(aev);
            max_idx = __shfl_sync(full, pos, last_a);
            cur_max = fmaxf(__shfl_sync(full, incl, 31), cur_max);
          }
          break;
        }
      }
      w.G[i0 + lane] = mres;
      cir += 32;
      if (cir >= A) cir -= A;
    }
    __syncwarp();
    // ---- phase 3 (:224-228): raw gain ----
    bool any_lim = false;
#pragma unroll 4
    for (int i = lane; i < 1024; i += 32) {
      const float mx = w.G[i];
      const float g = mx > thr ? __fdiv_rn(thr, mx) : 1.0f;
      any_lim |= g < 1.0f;
      w.G[i] = g;
    }
    any_lim = __any_sync(full, any_lim);
    __syncwarp();
    // ---- phase 4 (:230-249): attack / release smoothing; at rest (gain 1 everywhere) it is the identity ----
    float min_gain = 1.0f;
    const bool active = any_lim || psg != 1.0 || gm != 1.0f;
    const bool defer = active && p.gbuf != nullptr;
    if (active && !defer) {
      // 32 samples per chunk: one coalesced load, the values reach the (warp-uniform) recursion through shuffles that do not
      // depend on it, the results are collected in the owning lane and stored once — no shared-memory latency on the chain
#pragma unroll 1
      for (int cb = 0; cb < 1024; cb += 32) {
       const float gv = w.G[cb + lane];
       float gout = 0.0f;
#pragma unroll
       for (int q = 0; q < 32; q++) {
        const float go = pl_smooth_step(__shfl_sync(full, gv, q), gm, psg, (double)ac, (double)rc);
        gout = lane == q ? go : gout;
        if (go < min_gain) min_gain = go;
       }
       w.G[cb + lane] = gout;
      }
      __syncwarp();
    }
    // ---- max_buf holds the channel maxima of the last A samples of this frame ----
#pragma unroll 1
    for (int i = 1024 - A + lane; i < 1024; i += 32) mbg[(cir0 + i) % A] = w.T[i];
    if (defer) {  // raw gains to scratch, recursion and output in the two follow-up kernels
#pragma unroll 4
      for (int i = lane; i < 1024; i += 32) p.gbuf[u * 1024 + i] = w.G[i];
      if (lane == 0) {
        st[kPlCir] = cir;
        st[kPlMaxIdx] = max_idx;
        p.list[atomicAdd(p.count, 1)] = (int)u;
      }
      continue;
    }
    pl_emit_frame(p, u, st, w.G, lane, ch, A, di0, gt0, gt1);
    if (lane == 0) {
      st[kPlGainMod] = __float_as_int(gm);
      st[kPlMinGain] = __float_as_int(min_gain);
      st[kPlPsg] = __double2loint(psg);
      st[kPlPsg + 1] = __double2hiint(psg);
      st[kPlCir] = cir;
      st[kPlMaxIdx] = max_idx;
      st[kPlDelayIdx] = (di0 + 1024) % A;
      if (p.err) p.err[u] = 0;
    }
  }
}

// Attack / release recursion of the deferred streams, lane = stream: a warp takes 32 queued streams.  Each lane reads its own
// stream's gains 32 at a time as eight 16-byte requests (every request of the warp touches 32 different 128-byte lines, each one
// used in full — no transposition through shared memory, no shuffles), runs the recursion over them in registers with the next
// 32 already requested, and writes the smoothed gains back in place.  gain_modified, pre_smoothed_gain and min_gain go to the
// state records.
__global__ void __launch_bounds__(kPlSmoothWarps * 32) peak_limiter_smooth_kernel(PeakLimArgs p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = *p.count;
  const int groups = (n + 31) >> 5;
  for (int g = blockIdx.x * kPlSmoothWarps + warp; g < groups; g += gridDim.x * kPlSmoothWarps) {
    const int k = 32 * g + lane;
    if (k >= n) continue;
    const long long u = p.list[k];
    i32 *st = p.state + u * kPlWords;
    const double ac = (double)__int_as_float(st[kPlAttackConst]), rc = (double)__int_as_float(st[kPlReleaseConst]);
    float gm = __int_as_float(st[kPlGainMod]), min_gain = 1.f;
    double psg = __hiloint2double(st[kPlPsg + 1], st[kPlPsg]);
    float4 *gp = reinterpret_cast<float4 *>(p.gbuf + u * 1024);
    float4 nx[8];
#pragma unroll
    for (int q = 0; q < 8; q++) nx[q] = gp[q];
#pragma unroll 1
    for (int cb = 0; cb < 256; cb += 8) {
      float4 cur[8];
#pragma unroll
      for (int q = 0; q < 8; q++) cur[q] = nx[q];
      if (cb + 8 < 256) {
#pragma unroll
        for (int q = 0; q < 8; q++) nx[q] = gp[cb + 8 + q];
      }
#pragma unroll
      for (int q = 0; q < 8; q++) {
        float go;
        go = pl_smooth_step(cur[q].x, gm, psg, ac, rc); cur[q].x = go; min_gain = go < min_gain ? go : min_gain;
        go = pl_smooth_step(cur[q].y, gm, psg, ac, rc); cur[q].y = go; min_gain = go < min_gain ? go : min_gain;
        go = pl_smooth_step(cur[q].z, gm, psg, ac, rc); cur[q].z = go; min_gain = go < min_gain ? go : min_gain;
        go = pl_smooth_step(cur[q].w, gm, psg, ac, rc); cur[q].w = go; min_gain = go < min_gain ? go : min_gain;
      }
#pragma unroll
      for (int q = 0; q < 8; q++) gp[cb + q] = cur[q];
    }
    st[kPlGainMod] = __float_as_int(gm);
    st[kPlMinGain] = __float_as_int(min_gain);
    st[kPlPsg] = __double2loint(psg);
    st[kPlPsg + 1] = __double2hiint(psg);
  }
}

// Output of the deferred streams (phase 5) from the smoothed gains in scratch, one warp per stream.
__global__ void __launch_bounds__(kPlFinishWarps * 32) peak_limiter_finish_kernel(PeakLimArgs p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = *p.count;
  for (int k = blockIdx.x * kPlFinishWarps + warp; k < n; k += gridDim.x * kPlFinishWarps) {
    const long long u = p.list[k];
    i32 *st = p.state + u * kPlWords;
    const int ch = p.ch, A = st[kPlAttack], di0 = st[kPlDelayIdx];
    const float gt0 = (float)(1 << p.qshift_adj[u * ch]), gt1 = ch > 1 ? (float)(1 << p.qshift_adj[u * ch + 1]) : 0.f;
    pl_emit_frame(p, u, st, p.gbuf + u * 1024, lane, ch, A, di0, gt0, gt1);
    if (lane == 0) {
      st[kPlDelayIdx] = (di0 + 1024) % A;
      if (p.err) p.err[u] = 0;
    }
  }
}

size_t peak_limiter_scratch_bytes(long long n_units) { return (size_t)n_units * (4096 + 4) + 48; }

// scratch (peak_limiter_scratch_bytes, 16-byte aligned) or null: with scratch, streams whose smoothing recursion is active are
// finished by the two follow-up kernels.  which = 0 / 1 / 2 launches the main, the smoothing or the finishing kernel.
cudaError_t launch_peak_limiter(const PeakLimArgs &args_in, void *scratch, int which, int num_sms, cudaStream_t stream) {
  PeakLimArgs args = args_in;
  if (scratch) {
    args.count = reinterpret_cast<int *>(scratch);
    args.list = args.count + 4;
    args.gbuf = reinterpret_cast<float *>(args.list + args.n_units + (4 - (args.n_units & 3)) % 4);
  }
  if (which == 1) {
    long long need = ((args.n_units + 31) / 32 + kPlSmoothWarps - 1) / kPlSmoothWarps;
    long long grid = (long long)num_sms * 4;
    if (grid > need) grid = need;
    peak_limiter_smooth_kernel<<<(unsigned)(grid < 1 ? 1 : grid), kPlSmoothWarps * 32, 0, stream>>>(args);
    return cudaGetLastError();
  }
  long long need = (args.n_units + kPlWarps - 1) / kPlWarps;
  long long grid = (long long)num_sms * 5;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  if (which == 2) {  // no shared memory: 64 warps per SM
    long long g2 = (long long)num_sms * 8, need2 = (args.n_units + kPlFinishWarps - 1) / kPlFinishWarps;
    if (g2 > need2) g2 = need2;
    peak_limiter_finish_kernel<<<(unsigned)(g2 < 1 ? 1 : g2), kPlFinishWarps * 32, 0, stream>>>(args);
    return cudaGetLastError();
  }
  static xb::PerDeviceOnce configured;
  const size_t smem = sizeof(PlWarpS) * kPlWarps;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(peak_limiter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.done();
  }
  if (scratch) {
    cudaError_t e = cudaMemsetAsync(args.count, 0, 16, stream);
    if (e != cudaSuccess) return e;
  }
  peak_limiter_kernel<<<(unsigned)grid, kPlWarps * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace xb
