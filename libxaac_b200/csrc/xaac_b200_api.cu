// xaac_b200_api.cu — the C-ABI (include/xaac_b200.h) over the sm_100a kernels.
// No CPU fallback exists anywhere in this library: every entry point needs a CUDA device and reports
// XAAC_B200_ERR_CUDA otherwise.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <cuda_runtime.h>
#include "../../include/xaac_b200.h"
#include "kernels.h"

struct xaac_b200_ctx {
  int device = 0;
  int num_sms = 0;
  long long launches = 0;
  uint8_t *d_rom_imdct = nullptr;
  bool have_imdct_rom = false;
  uint8_t *d_rom_qmf_syn = nullptr;  // table image of qmf_synth_hq_kernel
  bool have_qmf_rom = false;
  int qmf_fast_bits = 0;
  alignas(16) uint8_t qmf_syn_tw[1024] = {0};
  cudaMemPool_t pool = nullptr;   // stream-ordered scratch of the stages that need some (kept across calls: no release threshold)  // twiddle image of qmf_synth_hq_kernel (kernel parameter, by value)
  uint8_t *d_rom_qmf_ana = nullptr;  // table image of qmf_anal_hq_kernel
  int qmf_anal_exact = 0;
  uint8_t *d_rom_lp = nullptr;       // table image of sbr_dec_lp_kernel (null: tables unsupported by the LP kernel)
  uint8_t *d_rom_env = nullptr;   // ia_env_calc_tables_struct
  uint8_t *d_rom_misc = nullptr;  // leading part of ixheaacd_misc_tables
  bool have_env_rom = false;
  uint8_t *d_rom_ps = nullptr;    // leading part of ia_ps_tables_struct
  uint8_t *d_rom_usac = nullptr;  // USAC FD tables (XAAC_UROM_*)
  uint8_t *d_rom_esbr = nullptr;  // table image of esbr_synth_kernel
  float *d_rom_rphase = nullptr;  // ixheaac_random_phase[512][2]
  float *d_rom_hbe = nullptr;     // XAAC_HROM_* blob of the harmonic transposer
  float *d_rom_fps = nullptr;     // XAAC_FPSROM_* blob of the float parametric stereo
  int32_t *d_rom_block = nullptr; // leading 620 bytes of ia_aac_dec_block_tables_struct (spectral stage)
  int esbr_periodic = 0;
  bool have_ps_rom = false;
  // XAAC_B200_DEV_CHUNK=N (experiment): the device-resident HQ stage runs in chunks of N units round-robin over dev_streams
  // internal streams, every stream re-using its own scratch slots so that the stage's matrices stay L2-resident
  long dev_chunk = 0;
  int dev_streams = 4;
  cudaStream_t dev_st[8] = {};
  cudaEvent_t dev_fork = nullptr, dev_join[8] = {};
  int sbr_unfused = 0;            // XAAC_B200_SBR_UNFUSED=1: the HQ stage launches its glue kernels separately (A/B, tests)
  int ps_rot_nosat = 0;           // no fractional-delay phase factor equals -32768 (16x16 rotations cannot saturate)
  char err[256] = {0};
  // optional per-kernel timing (CUDA events on the launching stream around every kernel)
  static constexpr int kMaxTicks = 8192;
  bool timing = false;
  int n_ticks = 0;
  const char *tick_name[kMaxTicks];
  cudaEvent_t tick_ev[kMaxTicks][2];
  // staging for the *_host entry points: kPipe chunks in flight, one stream each
#ifndef XB_PIPE
#define XB_PIPE 3
#endif
  static constexpr int kPipe = XB_PIPE;
  int64_t host_chunk = 4096;  // units per pipeline chunk of xaac_b200_heaac_frame_host (XAAC_B200_HOST_CHUNK overrides, tuning only)
  cudaStream_t streams[kPipe] = {};
  void *stage[kPipe] = {};
  size_t stage_bytes = 0;
};

struct xaac_b200_qmf_synth_state {
  int64_t n_units = 0;
  int16_t *d_states = nullptr;  // [n][1280]
  int16_t *d_pos = nullptr;     // [n][2]
};

struct xaac_b200_imdct_state {
  int64_t n_units = 0;
  int32_t *d_overlap = nullptr;  // [n][512]
  uint8_t *d_wstate = nullptr;   // [n][2] {window_shape, window_sequence}
};

// Device-resident state of the whole SBR stage, structure of arrays (one array per member of the host blob).
struct xaac_b200_sbr_state {
  int64_t n_units = 0;
  bool with_ps = false;
  bool lp_only = false;  // created for the low-power stage: no stage scratch
  // channel state (host blob XAAC_SBR_ST_*)
  int16_t *anal_states = nullptr, *anal_pos = nullptr, *syn_pos = nullptr, *sf = nullptr, *misc = nullptr, *env = nullptr,
          *syn_states = nullptr;
  int32_t *bw_prev = nullptr, *lpc = nullptr, *ov = nullptr;
  // PS state (host blob XAAC_PS_ST_*)
  int16_t *ps = nullptr, *syn_states_r = nullptr, *syn_pos_r = nullptr, *sf_r = nullptr;
  // stage scratch, sized by the units in flight (scratch_cap), not by the batch: a whole-batch _dev call needs n_units of it, the
  // chunked host pipeline kPipe chunks (ensure_sbr_scratch)
  int64_t scratch_cap = 0;
  int32_t *matrix = nullptr, *right = nullptr, *err = nullptr;
  int16_t *usb = nullptr, *hf_prm = nullptr, *synp = nullptr, *synp_r = nullptr, *ps_done = nullptr;
};

namespace {

struct BlobPart {
  void **dev;
  size_t bytes;   // per unit
  size_t offset;  // byte offset inside the host blob
};

void tick(xaac_b200_ctx *ctx, cudaStream_t st, const char *name, bool start) {
  if (!ctx->timing || ctx->n_ticks >= xaac_b200_ctx::kMaxTicks) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, st);
  if (start) {
    ctx->tick_name[ctx->n_ticks] = name;
    ctx->tick_ev[ctx->n_ticks][0] = ev;
  } else {
    ctx->tick_ev[ctx->n_ticks][1] = ev;
    ctx->n_ticks++;
  }
}

// a launch failed between the start and the stop tick: drop the start event
void tick_abort(xaac_b200_ctx *ctx) {
  if (!ctx->timing || ctx->n_ticks >= xaac_b200_ctx::kMaxTicks) return;
  cudaEventDestroy(ctx->tick_ev[ctx->n_ticks][0]);
}

// error exit of a *_host pipeline: earlier chunks may still have D2H copies into the caller's buffers in flight
int32_t drain(xaac_b200_ctx *ctx, int32_t rc) {
  if (rc != XAAC_B200_OK && ctx)
    for (int i = 0; i < xaac_b200_ctx::kPipe; i++)
      if (ctx->streams[i]) cudaStreamSynchronize(ctx->streams[i]);
  return rc;
}

int32_t fail(xaac_b200_ctx *ctx, cudaError_t e, const char *what) {
  if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, cudaGetErrorString(e));
  return XAAC_B200_ERR_CUDA;
}
int32_t bad_arg(xaac_b200_ctx *ctx, const char *what) {
  if (ctx) snprintf(ctx->err, sizeof(ctx->err), "bad argument: %s", what);
  return XAAC_B200_ERR_ARG;
}
// kernel launch with optional per-kernel CUDA-event timing (xaac_b200_kernel_timing)
#define LAUNCH(name, strm, call)                                   \
  do {                                                             \
    cudaError_t e__ = cudaSetDevice(ctx->device);                  \
    if (e__ != cudaSuccess) return fail(ctx, e__, "cudaSetDevice"); \
    tick(ctx, (cudaStream_t)(strm), name, true);                   \
    e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                      \
      tick_abort(ctx);                                             \
      return fail(ctx, e__, "launch " name);                       \
    }                                                              \
    tick(ctx, (cudaStream_t)(strm), name, false);                  \
  } while (0)
#define CK(call, what)                              \
  do {                                              \
    cudaError_t e__ = (call);                       \
    if (e__ != cudaSuccess) return fail(ctx, e__, what); \
  } while (0)

int32_t ensure_stage(xaac_b200_ctx *ctx, size_t bytes) {
  if (bytes <= ctx->stage_bytes) return XAAC_B200_OK;
  for (int i = 0; i < xaac_b200_ctx::kPipe; i++) {
    if (ctx->stage[i]) cudaFree(ctx->stage[i]);
    ctx->stage[i] = nullptr;
  }
  ctx->stage_bytes = 0;
  for (int i = 0; i < xaac_b200_ctx::kPipe; i++) CK(cudaMalloc(&ctx->stage[i], bytes), "cudaMalloc(stage)");
  ctx->stage_bytes = bytes;
  return XAAC_B200_OK;
}

}  // namespace

extern "C" {

int32_t xaac_b200_create(xaac_b200_ctx **out, int32_t device) {
  if (!out) return XAAC_B200_ERR_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    fprintf(stderr, "xaac_b200_create: no usable CUDA device %d (%s); this library has no CPU fallback\n", device,
            e == cudaSuccess ? "device index out of range" : cudaGetErrorString(e));
    return XAAC_B200_ERR_CUDA;
  }
  xaac_b200_ctx *ctx = new (std::nothrow) xaac_b200_ctx();
  if (!ctx) return XAAC_B200_FATAL;
  ctx->device = device;
  if (const char *uf = getenv("XAAC_B200_SBR_UNFUSED")) ctx->sbr_unfused = atoi(uf) != 0;
  if (const char *dc = getenv("XAAC_B200_DEV_CHUNK")) ctx->dev_chunk = atol(dc);
  if (const char *ds = getenv("XAAC_B200_DEV_STREAMS")) ctx->dev_streams = atoi(ds) < 1 ? 1 : (atoi(ds) > 8 ? 8 : atoi(ds));
  if (const char *hc = getenv("XAAC_B200_HOST_CHUNK")) {
    const long v = atol(hc);
    if (v >= 256 && v <= (1 << 20)) ctx->host_chunk = v;
  }
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess ||
      cudaMalloc((void **)&ctx->d_rom_imdct, xb::kRomImdctBytes + 64) != cudaSuccess) {
    delete ctx;
    return XAAC_B200_ERR_CUDA;
  }
  for (int i = 0; i < xaac_b200_ctx::kPipe; i++) {
    if (cudaStreamCreateWithFlags(&ctx->streams[i], cudaStreamNonBlocking) != cudaSuccess) {
      xaac_b200_destroy(ctx);
      return XAAC_B200_ERR_CUDA;
    }
  }
  *out = ctx;
  return XAAC_B200_OK;
}

void xaac_b200_destroy(xaac_b200_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (int i = 0; i < xaac_b200_ctx::kPipe; i++) {
    if (ctx->streams[i]) cudaStreamDestroy(ctx->streams[i]);
    if (ctx->stage[i]) cudaFree(ctx->stage[i]);
  }
  for (int k = 0; k < 8; k++) {  // streams / events of the opt-in chunked device path (XAAC_B200_DEV_CHUNK)
    if (ctx->dev_st[k]) cudaStreamDestroy(ctx->dev_st[k]);
    if (ctx->dev_join[k]) cudaEventDestroy(ctx->dev_join[k]);
  }
  if (ctx->dev_fork) cudaEventDestroy(ctx->dev_fork);
  if (ctx->d_rom_imdct) cudaFree(ctx->d_rom_imdct);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  if (ctx->d_rom_qmf_syn) cudaFree(ctx->d_rom_qmf_syn);
  if (ctx->d_rom_qmf_ana) cudaFree(ctx->d_rom_qmf_ana);
  if (ctx->d_rom_lp) cudaFree(ctx->d_rom_lp);
  if (ctx->d_rom_env) cudaFree(ctx->d_rom_env);
  if (ctx->d_rom_misc) cudaFree(ctx->d_rom_misc);
  if (ctx->d_rom_ps) cudaFree(ctx->d_rom_ps);
  if (ctx->d_rom_usac) cudaFree(ctx->d_rom_usac);
  if (ctx->d_rom_esbr) cudaFree(ctx->d_rom_esbr);
  if (ctx->d_rom_rphase) cudaFree(ctx->d_rom_rphase);
  if (ctx->d_rom_hbe) cudaFree(ctx->d_rom_hbe);
  if (ctx->d_rom_fps) cudaFree(ctx->d_rom_fps);
  if (ctx->d_rom_block) cudaFree(ctx->d_rom_block);
  delete ctx;
}

const char *xaac_b200_last_error(const xaac_b200_ctx *ctx) { return ctx ? ctx->err : "null context"; }
int32_t xaac_b200_num_sms(const xaac_b200_ctx *ctx) { return ctx ? ctx->num_sms : 0; }
int64_t xaac_b200_launch_count(const xaac_b200_ctx *ctx) { return ctx ? ctx->launches : 0; }

int32_t xaac_b200_sync(xaac_b200_ctx *ctx) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  CK(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  return XAAC_B200_OK;
}

int32_t xaac_b200_set_imdct_rom(xaac_b200_ctx *ctx, const void *tables, size_t bytes) {
  if (!ctx || !tables) return bad_arg(ctx, "null");
  if (bytes < (size_t)xb::kRomImdctBytes) return bad_arg(ctx, "IMDCT ROM blob shorter than 7500 bytes");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  // re-pack into the aligned device layout (kernels.h kDev*)
  uint8_t packed[xb::kDevImdctBytes];
  memset(packed, 0, sizeof(packed));
  const uint8_t *src = (const uint8_t *)tables;
  memcpy(packed + xb::kDevCos, src + xb::kRomCos, 1028);
  memcpy(packed + xb::kDevFftTw, src + xb::kRomFftTw, 1792);
  memcpy(packed + xb::kDevWinLongSine, src + xb::kRomWinLongSine, 2048);
  memcpy(packed + xb::kDevWinLongKbd, src + xb::kRomWinLongKbd, 2048);
  memcpy(packed + xb::kDevWinShortSine, src + xb::kRomWinShortSine, 256);
  memcpy(packed + xb::kDevWinShortKbd, src + xb::kRomWinShortKbd, 256);
  CK(cudaMemcpy(ctx->d_rom_imdct, packed, sizeof(packed), cudaMemcpyHostToDevice), "cudaMemcpy(rom)");
  ctx->have_imdct_rom = true;
  return XAAC_B200_OK;
}

int32_t xaac_b200_imdct_process_dev(xaac_b200_ctx *ctx, const int32_t *d_spec, int32_t *d_overlap,
                                    uint8_t *d_wstate, const uint8_t *d_ics, int32_t *d_out,
                                    int8_t *d_qshift_adj, int64_t n_units, int32_t ch_fac, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!ctx->have_imdct_rom) {
    snprintf(ctx->err, sizeof(ctx->err), "xaac_b200_set_imdct_rom has not been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (n_units < 0 || ch_fac < 1) return bad_arg(ctx, "n_units/ch_fac");
  if (n_units == 0) return XAAC_B200_OK;
  if (!d_spec || !d_overlap || !d_wstate || !d_ics || !d_out || !d_qshift_adj) return bad_arg(ctx, "null buffer");
  if (ch_fac > 1 && (n_units % ch_fac) != 0) return bad_arg(ctx, "n_units must be a multiple of ch_fac");
  xb::ImdctArgs a;
  a.spec = d_spec;
  a.overlap = d_overlap;
  a.wstate = d_wstate;
  a.ics = d_ics;
  a.out = d_out;
  a.qshift_adj = d_qshift_adj;
  a.rom = ctx->d_rom_imdct;
  a.n_units = n_units;
  a.ch_fac = ch_fac;
  LAUNCH("imdct_ola_kernel", stream, xb::launch_imdct(a, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return XAAC_B200_OK;
}

int32_t xaac_b200_imdct_state_create(xaac_b200_ctx *ctx, int64_t n_units, xaac_b200_imdct_state **out) {
  if (!ctx || !out || n_units < 0) return bad_arg(ctx, "state_create");
  *out = nullptr;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  xaac_b200_imdct_state *st = new (std::nothrow) xaac_b200_imdct_state();
  if (!st) return XAAC_B200_FATAL;
  st->n_units = n_units;
  size_t n = (size_t)(n_units > 0 ? n_units : 1);
  cudaError_t e = cudaMalloc((void **)&st->d_overlap, n * 2048);
  if (e == cudaSuccess) e = cudaMalloc((void **)&st->d_wstate, n * 2);
  if (e == cudaSuccess) e = cudaMemset(st->d_overlap, 0, n * 2048);
  if (e == cudaSuccess) e = cudaMemset(st->d_wstate, 0, n * 2);
  if (e != cudaSuccess) {
    xaac_b200_imdct_state_destroy(ctx, st);
    return fail(ctx, e, "imdct_state_create");
  }
  *out = st;
  CK(cudaStreamSynchronize(0), "state init sync");  // order legacy-stream memsets / pageable copies before the non-blocking pipeline streams
  return XAAC_B200_OK;
}

void xaac_b200_imdct_state_destroy(xaac_b200_ctx *ctx, xaac_b200_imdct_state *st) {
  if (!st) return;
  if (ctx) cudaSetDevice(ctx->device);
  if (st->d_overlap) cudaFree(st->d_overlap);
  if (st->d_wstate) cudaFree(st->d_wstate);
  delete st;
}

int32_t xaac_b200_imdct_state_upload(xaac_b200_ctx *ctx, xaac_b200_imdct_state *st, const int32_t *overlap,
                                     const uint8_t *wstate) {
  if (!ctx || !st || !overlap || !wstate) return bad_arg(ctx, "state_upload");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  CK(cudaMemcpy(st->d_overlap, overlap, (size_t)st->n_units * 2048, cudaMemcpyHostToDevice), "H2D overlap");
  CK(cudaMemcpy(st->d_wstate, wstate, (size_t)st->n_units * 2, cudaMemcpyHostToDevice), "H2D wstate");
  CK(cudaStreamSynchronize(0), "state init sync");  // order legacy-stream memsets / pageable copies before the non-blocking pipeline streams
  return XAAC_B200_OK;
}

int32_t xaac_b200_imdct_state_download(xaac_b200_ctx *ctx, xaac_b200_imdct_state *st, int32_t *overlap,
                                       uint8_t *wstate) {
  if (!ctx || !st || !overlap || !wstate) return bad_arg(ctx, "state_download");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  CK(cudaMemcpy(overlap, st->d_overlap, (size_t)st->n_units * 2048, cudaMemcpyDeviceToHost), "D2H overlap");
  CK(cudaMemcpy(wstate, st->d_wstate, (size_t)st->n_units * 2, cudaMemcpyDeviceToHost), "D2H wstate");
  return XAAC_B200_OK;
}

// Host-buffer entry point: chunks the batch and runs H2D / kernel / D2H of successive chunks on kPipe
// streams so PCIe and the SMs overlap. Pinned host memory makes the copies truly asynchronous.
// The overlap/window state never leaves HBM.
static int32_t xaac_b200_imdct_process_host_impl(xaac_b200_ctx *ctx, xaac_b200_imdct_state *state, const int32_t *spec,
                                     const uint8_t *ics, int32_t *out, int8_t *qshift_adj, int32_t ch_fac) {
  if (!ctx || !state) return XAAC_B200_ERR_ARG;
  const int64_t n_units = state->n_units;
  if (ch_fac < 1) return bad_arg(ctx, "ch_fac");
  if (n_units == 0) return XAAC_B200_OK;
  if (!spec || !ics || !out || !qshift_adj) return bad_arg(ctx, "null buffer");
  if (ch_fac > 1 && (n_units % ch_fac) != 0) return bad_arg(ctx, "n_units must be a multiple of ch_fac");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  int64_t chunk = 8192;
  if (chunk > n_units) chunk = n_units;
  if (ch_fac > 1) chunk = ((chunk + ch_fac - 1) / ch_fac) * ch_fac;
  // per-unit staging layout: spec 4096 | out 4096 | ics 2 | qadj 1 (+pad)
  const size_t per_unit = 4096 + 4096 + 16;
  int32_t rc = ensure_stage(ctx, per_unit * (size_t)chunk);
  if (rc != XAAC_B200_OK) return rc;
  int slot = 0;
  for (int64_t u0 = 0; u0 < n_units; u0 += chunk, slot = (slot + 1) % xaac_b200_ctx::kPipe) {
    int64_t n = (n_units - u0 < chunk) ? (n_units - u0) : chunk;
    cudaStream_t st = ctx->streams[slot];
    uint8_t *base = (uint8_t *)ctx->stage[slot];
    int32_t *d_spec = (int32_t *)base;
    int32_t *d_out = (int32_t *)(base + 4096 * (size_t)chunk);
    uint8_t *d_ics = base + 8192 * (size_t)chunk;
    int8_t *d_qa = (int8_t *)(d_ics + 4 * (size_t)chunk);
    CK(cudaMemcpyAsync(d_spec, spec + u0 * 1024, (size_t)n * 4096, cudaMemcpyHostToDevice, st), "H2D spec");
    CK(cudaMemcpyAsync(d_ics, ics + u0 * 2, (size_t)n * 2, cudaMemcpyHostToDevice, st), "H2D ics");
    rc = xaac_b200_imdct_process_dev(ctx, d_spec, state->d_overlap + u0 * 512, state->d_wstate + u0 * 2, d_ics,
                                     d_out, d_qa, n, ch_fac, st);
    if (rc != XAAC_B200_OK) return rc;
    CK(cudaMemcpyAsync(out + u0 * 1024, d_out, (size_t)n * 4096, cudaMemcpyDeviceToHost, st), "D2H out");
    CK(cudaMemcpyAsync(qshift_adj + u0, d_qa, (size_t)n, cudaMemcpyDeviceToHost, st), "D2H qshift_adj");
  }
  for (int i = 0; i < xaac_b200_ctx::kPipe; i++) CK(cudaStreamSynchronize(ctx->streams[i]), "stream sync");
  return XAAC_B200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// QMF banks
// ---------------------------------------------------------------------------------------------------------
int32_t xaac_b200_set_qmf_rom(xaac_b200_ctx *ctx, const void *tables, size_t bytes) {
  if (!ctx || !tables) return bad_arg(ctx, "null");
  if (bytes < (size_t)xb::kQRomBytes) return bad_arg(ctx, "QMF ROM blob shorter than 3464 bytes");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  size_t n = xb::qmf_synth_table_bytes();
  uint8_t *img = (uint8_t *)calloc(1, n + 64);
  if (!img) return XAAC_B200_FATAL;
  int fast_bits = xb::qmf_synth_build_tables((const uint8_t *)tables, img);
  if (fast_bits <= 0) {
    free(img);
    return bad_arg(ctx, "QMF tables: unexpected digit-reverse table or prototype filter exceeds the no-saturation bound");
  }
  if (!ctx->d_rom_qmf_syn) {
    cudaError_t e = cudaMalloc((void **)&ctx->d_rom_qmf_syn, n);
    if (e != cudaSuccess) {
      free(img);
      return fail(ctx, e, "cudaMalloc(qmf rom)");
    }
  }
  cudaError_t e = cudaMemcpy(ctx->d_rom_qmf_syn, img, n, cudaMemcpyHostToDevice);
  free(img);
  if (e != cudaSuccess) return fail(ctx, e, "cudaMemcpy(qmf rom)");
  {  // analysis bank tables
    size_t na = xb::qmf_anal_table_bytes();
    uint8_t *ia = (uint8_t *)calloc(1, na + 64);
    if (!ia) return XAAC_B200_FATAL;
    int ex = xb::qmf_anal_build_tables((const uint8_t *)tables, ia);
    if (ex < 0) {
      free(ia);
      return bad_arg(ctx, "QMF tables: unexpected dig_rev_table4_16");
    }
    if (!ctx->d_rom_qmf_ana) {
      cudaError_t e2 = cudaMalloc((void **)&ctx->d_rom_qmf_ana, na);
      if (e2 != cudaSuccess) {
        free(ia);
        return fail(ctx, e2, "cudaMalloc(qmf anal rom)");
      }
    }
    cudaError_t e2 = cudaMemcpy(ctx->d_rom_qmf_ana, ia, na, cudaMemcpyHostToDevice);
    free(ia);
    if (e2 != cudaSuccess) return fail(ctx, e2, "cudaMemcpy(qmf anal rom)");
    ctx->qmf_anal_exact = ex;
  }
  {  // low-power stage tables (dct3_32 / dct2_64 twiddles, prototype)
    size_t nl = xb::sbr_lp_table_bytes();
    uint8_t *il = (uint8_t *)calloc(1, nl + 64);
    if (!il) return XAAC_B200_FATAL;
    if (xb::sbr_lp_build_tables((const uint8_t *)tables, il) == 0) {
      cudaError_t e3 = ctx->d_rom_lp ? cudaSuccess : cudaMalloc((void **)&ctx->d_rom_lp, nl);
      if (e3 == cudaSuccess) e3 = cudaMemcpy(ctx->d_rom_lp, il, nl, cudaMemcpyHostToDevice);
      free(il);
      if (e3 != cudaSuccess) return fail(ctx, e3, "cudaMemcpy(lp rom)");
    } else {
      free(il);
      if (ctx->d_rom_lp) cudaFree(ctx->d_rom_lp);
      ctx->d_rom_lp = nullptr;
    }
  }
  if (xb::qmf_synth_twiddle_bytes() > sizeof(ctx->qmf_syn_tw)) return XAAC_B200_FATAL;
  xb::qmf_synth_build_twiddles((const uint8_t *)tables, ctx->qmf_syn_tw);
  ctx->have_qmf_rom = true;
  ctx->qmf_fast_bits = fast_bits;
  return XAAC_B200_OK;
}

int32_t xaac_b200_qmf_synth_hq_dev(xaac_b200_ctx *ctx, const int32_t *d_matrix, int16_t *d_filter_states,
                                   int16_t *d_pos, const int16_t *d_params, int16_t *d_pcm, int64_t n_units,
                                   int32_t ch_fac, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!ctx->have_qmf_rom) {
    snprintf(ctx->err, sizeof(ctx->err), "xaac_b200_set_qmf_rom has not been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (n_units < 0 || ch_fac < 1) return bad_arg(ctx, "n_units/ch_fac");
  if (n_units == 0) return XAAC_B200_OK;
  if (!d_matrix || !d_filter_states || !d_pos || !d_params || !d_pcm) return bad_arg(ctx, "null buffer");
  if (ch_fac > 1 && (n_units % ch_fac) != 0) return bad_arg(ctx, "n_units must be a multiple of ch_fac");
  xb::QmfSynthArgs a;
  a.matrix = d_matrix;
  a.states = d_filter_states;
  a.pos = d_pos;
  a.params = d_params;
  a.pcm = d_pcm;
  a.rom = ctx->d_rom_qmf_syn;
  a.twiddles = ctx->qmf_syn_tw;
  a.fast_bits = ctx->qmf_fast_bits;
  a.zero = 0;
  a.n_units = n_units;
  a.ch_fac = ch_fac;
  LAUNCH("qmf_synth_hq_kernel", stream, xb::launch_qmf_synth_hq(a, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return XAAC_B200_OK;
}

int32_t xaac_b200_qmf_synth_state_create(xaac_b200_ctx *ctx, int64_t n_units, xaac_b200_qmf_synth_state **out) {
  if (!ctx || !out || n_units < 0) return bad_arg(ctx, "state_create");
  *out = nullptr;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  xaac_b200_qmf_synth_state *st = new (std::nothrow) xaac_b200_qmf_synth_state();
  if (!st) return XAAC_B200_FATAL;
  st->n_units = n_units;
  size_t n = (size_t)(n_units > 0 ? n_units : 1);
  cudaError_t e = cudaMalloc((void **)&st->d_states, n * 2560);
  if (e == cudaSuccess) e = cudaMalloc((void **)&st->d_pos, n * 4);
  if (e == cudaSuccess) e = cudaMemset(st->d_states, 0, n * 2560);
  if (e == cudaSuccess) e = cudaMemset(st->d_pos, 0, n * 4);
  if (e != cudaSuccess) {
    xaac_b200_qmf_synth_state_destroy(ctx, st);
    return fail(ctx, e, "qmf_synth_state_create");
  }
  *out = st;
  CK(cudaStreamSynchronize(0), "state init sync");  // order legacy-stream memsets / pageable copies before the non-blocking pipeline streams
  return XAAC_B200_OK;
}

void xaac_b200_qmf_synth_state_destroy(xaac_b200_ctx *ctx, xaac_b200_qmf_synth_state *st) {
  if (!st) return;
  if (ctx) cudaSetDevice(ctx->device);
  if (st->d_states) cudaFree(st->d_states);
  if (st->d_pos) cudaFree(st->d_pos);
  delete st;
}

int32_t xaac_b200_qmf_synth_state_upload(xaac_b200_ctx *ctx, xaac_b200_qmf_synth_state *st,
                                         const int16_t *filter_states, const int16_t *pos) {
  if (!ctx || !st || !filter_states || !pos) return bad_arg(ctx, "state_upload");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  CK(cudaMemcpy(st->d_states, filter_states, (size_t)st->n_units * 2560, cudaMemcpyHostToDevice), "H2D states");
  CK(cudaMemcpy(st->d_pos, pos, (size_t)st->n_units * 4, cudaMemcpyHostToDevice), "H2D pos");
  CK(cudaStreamSynchronize(0), "state init sync");  // order legacy-stream memsets / pageable copies before the non-blocking pipeline streams
  return XAAC_B200_OK;
}

int32_t xaac_b200_qmf_synth_state_download(xaac_b200_ctx *ctx, xaac_b200_qmf_synth_state *st,
                                           int16_t *filter_states, int16_t *pos) {
  if (!ctx || !st || !filter_states || !pos) return bad_arg(ctx, "state_download");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  CK(cudaMemcpy(filter_states, st->d_states, (size_t)st->n_units * 2560, cudaMemcpyDeviceToHost), "D2H states");
  CK(cudaMemcpy(pos, st->d_pos, (size_t)st->n_units * 4, cudaMemcpyDeviceToHost), "D2H pos");
  return XAAC_B200_OK;
}

static int32_t xaac_b200_qmf_synth_hq_host_impl(xaac_b200_ctx *ctx, xaac_b200_qmf_synth_state *state, const int32_t *matrix,
                                    const int16_t *params, int16_t *pcm, int32_t ch_fac) {
  if (!ctx || !state) return XAAC_B200_ERR_ARG;
  const int64_t n_units = state->n_units;
  if (ch_fac < 1) return bad_arg(ctx, "ch_fac");
  if (n_units == 0) return XAAC_B200_OK;
  if (!matrix || !params || !pcm) return bad_arg(ctx, "null buffer");
  if (ch_fac > 1 && (n_units % ch_fac) != 0) return bad_arg(ctx, "n_units must be a multiple of ch_fac");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  int64_t chunk = 4096;
  if (chunk > n_units) chunk = n_units;
  if (ch_fac > 1) chunk = ((chunk + ch_fac - 1) / ch_fac) * ch_fac;
  // per-unit staging layout: matrix 16384 | pcm 4096 | params 16
  const size_t per_unit = 16384 + 4096 + 16;
  int32_t rc = ensure_stage(ctx, per_unit * (size_t)chunk);
  if (rc != XAAC_B200_OK) return rc;
  int slot = 0;
  for (int64_t u0 = 0; u0 < n_units; u0 += chunk, slot = (slot + 1) % xaac_b200_ctx::kPipe) {
    int64_t n = (n_units - u0 < chunk) ? (n_units - u0) : chunk;
    cudaStream_t st = ctx->streams[slot];
    uint8_t *base = (uint8_t *)ctx->stage[slot];
    int32_t *d_mat = (int32_t *)base;
    int16_t *d_pcm = (int16_t *)(base + 16384 * (size_t)chunk);
    int16_t *d_prm = (int16_t *)(base + 20480 * (size_t)chunk);
    CK(cudaMemcpyAsync(d_mat, matrix + u0 * 4096, (size_t)n * 16384, cudaMemcpyHostToDevice, st), "H2D matrix");
    CK(cudaMemcpyAsync(d_prm, params + u0 * 8, (size_t)n * 16, cudaMemcpyHostToDevice, st), "H2D params");
    rc = xaac_b200_qmf_synth_hq_dev(ctx, d_mat, state->d_states + u0 * 1280, state->d_pos + u0 * 2, d_prm, d_pcm, n,
                                    ch_fac, st);
    if (rc != XAAC_B200_OK) return rc;
    CK(cudaMemcpyAsync(pcm + u0 * 2048, d_pcm, (size_t)n * 4096, cudaMemcpyDeviceToHost, st), "D2H pcm");
  }
  for (int i = 0; i < xaac_b200_ctx::kPipe; i++) CK(cudaStreamSynchronize(ctx->streams[i]), "stream sync");
  return XAAC_B200_OK;
}

int32_t xaac_b200_qmf_anal_hq_dev(xaac_b200_ctx *ctx, const int16_t *d_pcm, int16_t *d_states, int16_t *d_pos,
                                  const int16_t *d_usb, int32_t *d_matrix, int64_t n_units, int32_t ch_fac,
                                  void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!ctx->have_qmf_rom) {
    snprintf(ctx->err, sizeof(ctx->err), "xaac_b200_set_qmf_rom has not been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (n_units < 0 || ch_fac < 1) return bad_arg(ctx, "n_units/ch_fac");
  if (n_units == 0) return XAAC_B200_OK;
  if (!d_pcm || !d_states || !d_pos || !d_usb || !d_matrix) return bad_arg(ctx, "null buffer");
  if (ch_fac > 1 && (n_units % ch_fac) != 0) return bad_arg(ctx, "n_units must be a multiple of ch_fac");
  xb::QmfAnalArgs a;
  a.pcm = d_pcm;
  a.states = d_states;
  a.pos = d_pos;
  a.usb = d_usb;
  a.matrix = d_matrix;
  a.rom = ctx->d_rom_qmf_ana;
  a.n_units = n_units;
  a.ch_fac = ch_fac;
  a.exact = ctx->qmf_anal_exact;
  LAUNCH("qmf_anal_hq_kernel", stream, xb::launch_qmf_anal_hq(a, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return XAAC_B200_OK;
}

int32_t xaac_b200_hf_generator_hq_dev(xaac_b200_ctx *ctx, const int32_t *d_lpc, int32_t *d_matrix,
                                      const int16_t *d_params, int32_t *d_bw_prev, int16_t *d_hb_scale,
                                      int64_t n_units, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (n_units < 0) return bad_arg(ctx, "n_units");
  if (n_units == 0) return XAAC_B200_OK;
  if (!d_lpc || !d_matrix || !d_params || !d_bw_prev || !d_hb_scale) return bad_arg(ctx, "null buffer");
  xb::HfGenArgs a;
  a.lpc = d_lpc;
  a.matrix = d_matrix;
  a.params = d_params;
  a.bw_prev = d_bw_prev;
  a.hb_scale = d_hb_scale;
  a.n_units = n_units;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  LAUNCH("hf_generator_hq_kernel", stream, xb::launch_hf_generator_hq(a, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return XAAC_B200_OK;
}

int32_t xaac_b200_set_env_rom(xaac_b200_ctx *ctx, const void *env_tables, size_t env_bytes, const void *misc_tables,
                              size_t misc_bytes) {
  if (!ctx || !env_tables || !misc_tables) return bad_arg(ctx, "null");
  if (env_bytes < (size_t)xb::kERomBytes) return bad_arg(ctx, "env ROM blob shorter than 2404 bytes");
  if (misc_bytes < (size_t)xb::kMRomBytes) return bad_arg(ctx, "misc ROM blob shorter than 2470 bytes");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  if (!ctx->d_rom_env) CK(cudaMalloc((void **)&ctx->d_rom_env, xb::kERomBytes + 60), "cudaMalloc(env rom)");
  if (!ctx->d_rom_misc) CK(cudaMalloc((void **)&ctx->d_rom_misc, xb::kMRomBytes + 58), "cudaMalloc(misc rom)");
  CK(cudaMemcpy(ctx->d_rom_env, env_tables, xb::kERomBytes, cudaMemcpyHostToDevice), "H2D env rom");
  CK(cudaMemcpy(ctx->d_rom_misc, misc_tables, xb::kMRomBytes, cudaMemcpyHostToDevice), "H2D misc rom");
  ctx->have_env_rom = true;
  return XAAC_B200_OK;
}

int32_t xaac_b200_calc_sbrenvelope_hq_dev(xaac_b200_ctx *ctx, const int16_t *d_params, int16_t *d_sf, int16_t *d_state,
                                          int32_t *d_matrix, int32_t *d_err, int64_t n_units, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!ctx->have_env_rom) {
    snprintf(ctx->err, sizeof(ctx->err), "xaac_b200_set_env_rom has not been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (n_units < 0) return bad_arg(ctx, "n_units");
  if (n_units == 0) return XAAC_B200_OK;
  if (!d_params || !d_sf || !d_state || !d_matrix) return bad_arg(ctx, "null buffer");
  xb::EnvCalcArgs a;
  a.params = d_params;
  a.sf = d_sf;
  a.state = d_state;
  a.matrix = d_matrix;
  a.err = d_err;
  a.env_rom = ctx->d_rom_env;
  a.misc_rom = ctx->d_rom_misc;
  a.n_units = n_units;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  LAUNCH("calc_sbrenvelope_hq_kernel", stream, xb::launch_calc_sbrenvelope_hq(a, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return XAAC_B200_OK;
}

int32_t xaac_b200_set_ps_rom(xaac_b200_ctx *ctx, const void *ps_tables, size_t bytes) {
  if (!ctx || !ps_tables) return bad_arg(ctx, "null");
  if (bytes < (size_t)xb::kPsRomBytes) return bad_arg(ctx, "PS ROM blob shorter than 1230 bytes");
  {  // the PS kernel has the QMF band -> parameter bin map of borders_group[10..22] compiled in (power pre-pass)
    static const int16_t kStdBorders[13] = {3, 4, 5, 6, 7, 8, 9, 11, 14, 18, 23, 35, 64};
    const int16_t *t = (const int16_t *)ps_tables + xb::kPsRomBordersGroup + 10;
    for (int i = 0; i < 13; i++)
      if (t[i] != kStdBorders[i]) return bad_arg(ctx, "PS tables: unexpected borders_group");
  }
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  if (!ctx->d_rom_ps) CK(cudaMalloc((void **)&ctx->d_rom_ps, xb::kPsRomBytes + 50), "cudaMalloc(ps rom)");
  CK(cudaMemcpy(ctx->d_rom_ps, ps_tables, xb::kPsRomBytes, cudaMemcpyHostToDevice), "H2D ps rom");
  {
    const int16_t *t = (const int16_t *)ps_tables;
    int ok = 1;
    for (int i = xb::kPsRomFracQmf; i < xb::kPsRomScale; i++)
      if (t[i] == -32768) ok = 0;
    ctx->ps_rot_nosat = ok;
  }
  ctx->have_ps_rom = true;
  return XAAC_B200_OK;
}

static void sbr_state_parts(xaac_b200_sbr_state *s, BlobPart *st, int *n_st, BlobPart *ps, int *n_ps) {
  int i = 0;
  st[i++] = {(void **)&s->anal_states, 640, 2 * (size_t)xb::kSbrStAnalStates};
  st[i++] = {(void **)&s->anal_pos, 4, 2 * (size_t)xb::kSbrStAnalPos};
  st[i++] = {(void **)&s->syn_pos, 4, 2 * (size_t)xb::kSbrStSynPos};
  st[i++] = {(void **)&s->sf, 16, 2 * (size_t)xb::kSbrStSf};
  st[i++] = {(void **)&s->misc, 32, 2 * (size_t)xb::kSbrStMisc};
  st[i++] = {(void **)&s->env, 464, 2 * (size_t)xb::kSbrStEnv};
  st[i++] = {(void **)&s->syn_states, 2560, 2 * (size_t)xb::kSbrStSynStates};
  st[i++] = {(void **)&s->bw_prev, 24, 2 * (size_t)xb::kSbrStBwPrev};
  st[i++] = {(void **)&s->lpc, 1024, 2 * (size_t)xb::kSbrStLpc};
  st[i++] = {(void **)&s->ov, 3072, 2 * (size_t)xb::kSbrStOv};
  *n_st = i;
  i = 0;
  ps[i++] = {(void **)&s->ps, 2 * (size_t)xb::kPsDspWords, 0};
  ps[i++] = {(void **)&s->syn_states_r, 2560, 2 * (size_t)xb::kPsStSynStatesR};
  ps[i++] = {(void **)&s->syn_pos_r, 4, 2 * (size_t)xb::kPsStSynPosR};
  ps[i++] = {(void **)&s->sf_r, 16, 2 * (size_t)xb::kPsStSfR};
  *n_ps = i;
}

void xaac_b200_sbr_state_destroy(xaac_b200_ctx *ctx, xaac_b200_sbr_state *s) {
  if (!s) return;
  if (ctx) cudaSetDevice(ctx->device);
  BlobPart st[16], ps[8];
  int n_st, n_ps;
  sbr_state_parts(s, st, &n_st, ps, &n_ps);
  for (int i = 0; i < n_st; i++) if (*st[i].dev) cudaFree(*st[i].dev);
  for (int i = 0; i < n_ps; i++) if (*ps[i].dev) cudaFree(*ps[i].dev);
  void *scr[] = {s->matrix, s->right, s->err, s->usb, s->hf_prm, s->synp, s->synp_r, s->ps_done};
  for (void *q : scr) if (q) cudaFree(q);
  delete s;
}

int32_t xaac_b200_sbr_state_create(xaac_b200_ctx *ctx, int64_t n_units, int32_t with_ps, xaac_b200_sbr_state **out) {
  if (!ctx || !out || n_units <= 0) return bad_arg(ctx, "sbr_state_create");
  *out = nullptr;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  xaac_b200_sbr_state *s = new (std::nothrow) xaac_b200_sbr_state();
  if (!s) return XAAC_B200_FATAL;
  s->n_units = n_units;
  s->lp_only = with_ps == XAAC_B200_SBR_STATE_LP;
  s->with_ps = with_ps != 0 && !s->lp_only;
  BlobPart st[16], ps[8];
  int n_st, n_ps;
  sbr_state_parts(s, st, &n_st, ps, &n_ps);
  bool ok = true;
  auto alloc0 = [&](void **q, size_t bytes) {
    if (!ok) return;
    if (cudaMalloc(q, bytes) != cudaSuccess || cudaMemset(*q, 0, bytes) != cudaSuccess) ok = false;
  };
  for (int i = 0; i < n_st; i++) alloc0(st[i].dev, st[i].bytes * (size_t)n_units);
  if (s->with_ps)
    for (int i = 0; i < n_ps; i++) alloc0(ps[i].dev, ps[i].bytes * (size_t)n_units);
  alloc0((void **)&s->err, (size_t)n_units * 4);
  if (!ok) {
    cudaError_t e = cudaGetLastError();
    xaac_b200_sbr_state_destroy(ctx, s);
    return fail(ctx, e, "cudaMalloc(sbr state)");
  }
  *out = s;
  CK(cudaStreamSynchronize(0), "state init sync");  // order legacy-stream memsets / pageable copies before the non-blocking pipeline streams
  return XAAC_B200_OK;
}

static int32_t sbr_state_copy(xaac_b200_ctx *ctx, xaac_b200_sbr_state *s, int16_t *st_blob, int16_t *ps_blob, bool up) {
  if (!ctx || !s) return XAAC_B200_ERR_ARG;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  BlobPart st[16], ps[8];
  int n_st, n_ps;
  sbr_state_parts(s, st, &n_st, ps, &n_ps);
  auto run = [&](BlobPart *parts, int n, int16_t *blob, size_t pitch) -> cudaError_t {
    for (int i = 0; i < n; i++) {
      uint8_t *h = (uint8_t *)blob + parts[i].offset;
      cudaError_t e = up ? cudaMemcpy2D(*parts[i].dev, parts[i].bytes, h, pitch, parts[i].bytes, (size_t)s->n_units,
                                        cudaMemcpyHostToDevice)
                         : cudaMemcpy2D(h, pitch, *parts[i].dev, parts[i].bytes, parts[i].bytes, (size_t)s->n_units,
                                        cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  };
  if (st_blob) CK(run(st, n_st, st_blob, 2 * (size_t)xb::kSbrStWords), "sbr state copy");
  if (ps_blob) {
    if (!s->with_ps) return bad_arg(ctx, "state was created without PS");
    CK(run(ps, n_ps, ps_blob, 2 * (size_t)xb::kPsStWords), "ps state copy");
  }
  CK(cudaStreamSynchronize(0), "state copy sync");  // order against the non-blocking pipeline streams
  return XAAC_B200_OK;
}

int32_t xaac_b200_sbr_state_upload(xaac_b200_ctx *ctx, xaac_b200_sbr_state *s, const int16_t *st_blob,
                                   const int16_t *ps_blob) {
  return sbr_state_copy(ctx, s, (int16_t *)st_blob, (int16_t *)ps_blob, true);
}
int32_t xaac_b200_sbr_state_download(xaac_b200_ctx *ctx, xaac_b200_sbr_state *s, int16_t *st_blob, int16_t *ps_blob) {
  return sbr_state_copy(ctx, s, st_blob, ps_blob, false);
}

extern "C" int32_t xaac_b200_imdct_out_to_pcm16_dev(xaac_b200_ctx *, const int32_t *, const int8_t *, int16_t *, int64_t,
                                                    int32_t, void *);
// (re)allocate the stage scratch of an HQ state for `units` units in flight (synchronises the device when it grows)
static int32_t ensure_sbr_scratch(xaac_b200_ctx *ctx, xaac_b200_sbr_state *s, int64_t units) {
  if (s->lp_only || units <= s->scratch_cap) return XAAC_B200_OK;
  CK(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  void **old[] = {(void **)&s->matrix, (void **)&s->right, (void **)&s->usb, (void **)&s->hf_prm, (void **)&s->synp,
                  (void **)&s->synp_r, (void **)&s->ps_done};
  for (void **q : old) {
    if (*q) cudaFree(*q);
    *q = nullptr;
  }
  s->scratch_cap = 0;
  bool ok = true;
  auto alloc0 = [&](void **q, size_t bytes) {
    if (!ok) return;
    if (cudaMalloc(q, bytes) != cudaSuccess || cudaMemset(*q, 0, bytes) != cudaSuccess) ok = false;
  };
  alloc0((void **)&s->matrix, (size_t)units * xb::kSbrMatWords * 4);
  alloc0((void **)&s->usb, (size_t)units * 2);
  alloc0((void **)&s->hf_prm, (size_t)units * 160);
  alloc0((void **)&s->synp, (size_t)units * 16);
  if (s->with_ps) {
    alloc0((void **)&s->right, (size_t)units * 4096 * 4);
    alloc0((void **)&s->synp_r, (size_t)units * 16);
    alloc0((void **)&s->ps_done, (size_t)units * 2);
  }
  if (!ok) return fail(ctx, cudaGetLastError(), "cudaMalloc(sbr stage scratch)");
  CK(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  s->scratch_cap = units;
  return XAAC_B200_OK;
}
// One frame for units [u0, u0 + n) of the state, using scratch slots [su0, su0 + n); d_side / d_time_in / d_time_out / d_err
// point at the chunk's first unit.
// d_w32 != null: the core coder's WORD32 output + qshift_adj of the chunk instead of d_time_in (converted on load).
static int32_t sbr_dec_range(xaac_b200_ctx *ctx, xaac_b200_sbr_state *s, long long u0, long long n, const int16_t *d_side,
                             const int16_t *d_time_in, int16_t *d_time_out, int32_t *d_err, cudaStream_t st, long long su0,
                             const int32_t *d_w32 = nullptr, const int8_t *d_adj = nullptr) {
  int32_t *err = d_err ? d_err : s->err + u0;
  xb::SbrStageArgs g;
  g.side = d_side; g.matrix = s->matrix + su0 * xb::kSbrMatWords; g.ov = s->ov + u0 * 768; g.lpc = s->lpc + u0 * 256;
  g.sf = s->sf + u0 * 8; g.misc = s->misc + u0 * 16; g.usb = s->usb + su0; g.hf_prm = s->hf_prm + su0 * 80;
  g.synp = s->synp + su0 * 8; g.err = err; g.n_units = n;
  {
    xb::QmfAnalArgs a;
    a.pcm = d_time_in; a.states = s->anal_states + u0 * 320; a.pos = s->anal_pos + u0 * 2; a.usb = g.usb;
    a.matrix = g.matrix + 6 * 128; a.rom = ctx->d_rom_qmf_ana; a.n_units = n; a.ch_fac = 1;
    a.exact = ctx->qmf_anal_exact; a.mat_stride = xb::kSbrMatWords; a.w32 = d_w32; a.qshift_adj = d_adj;
    if (ctx->sbr_unfused) {
      LAUNCH("sbr_pre_kernel", st, xb::launch_sbr_pre(g, ctx->num_sms, st));
      LAUNCH("qmf_anal_hq_kernel", st, xb::launch_qmf_anal_hq(a, ctx->num_sms, st));
      LAUNCH("sbr_scale_kernel", st, xb::launch_sbr_scale(g, ctx->num_sms, st));
      ctx->launches += 2;
    } else {  // the three of them per unit by one warp
      LAUNCH("sbr_front_hq_kernel", st, xb::launch_sbr_front_hq(a, g, ctx->num_sms, st));
    }
  }
  {
    xb::HfGenArgs a;
    a.lpc = g.lpc; a.matrix = g.matrix; a.params = g.hf_prm; a.bw_prev = s->bw_prev + u0 * 6;
    a.hb_scale = g.sf + xb::kSfHb; a.hb_stride = 8; a.n_units = n;
    a.gate = d_side + xb::kSideApply; a.gate_stride = xb::kSideWords;
    LAUNCH("hf_generator_hq_kernel", st, xb::launch_hf_generator_hq(a, ctx->num_sms, st));
  }
  {
    xb::EnvCalcArgs a;
    a.params = d_side; a.prm_stride = xb::kSideWords; a.sf = g.sf; a.state = s->env + u0 * xb::kEnvStWords;
    a.matrix = g.matrix; a.err = err; a.env_rom = ctx->d_rom_env; a.misc_rom = ctx->d_rom_misc; a.n_units = n;
    a.gate = d_side + xb::kSideApply; a.gate_stride = xb::kSideWords; a.max_qmf_prev = g.misc;
    if (ctx->sbr_unfused) {
      LAUNCH("calc_sbrenvelope_hq_kernel", st, xb::launch_calc_sbrenvelope_hq(a, ctx->num_sms, st));
      LAUNCH("sbr_post_kernel", st, xb::launch_sbr_post(g, ctx->num_sms, st));
      ctx->launches += 1;
    } else {
      LAUNCH("calc_sbrenvelope_hq_post_kernel", st, xb::launch_calc_sbrenvelope_hq_post(a, g, ctx->num_sms, st));
    }
  }
  ctx->launches += 3;
  xb::QmfSynthArgs y;
  y.matrix = g.matrix; y.states = s->syn_states + u0 * 1280; y.pos = s->syn_pos + u0 * 2; y.params = g.synp;
  y.pcm = d_time_out; y.rom = ctx->d_rom_qmf_syn; y.twiddles = ctx->qmf_syn_tw; y.n_units = n; y.fast_bits = ctx->qmf_fast_bits; y.zero = 0;
  y.mat_stride = xb::kSbrMatWords;
  if (s->with_ps) {
    xb::PsArgs a;
    a.side = d_side; a.matrix = g.matrix; a.right = s->right + su0 * 4096; a.ps_state = s->ps + u0 * xb::kPsDspWords;
    a.sf = g.sf; a.sf_r = s->sf_r + u0 * 8; a.synp = g.synp; a.synp_r = s->synp_r + su0 * 8; a.ps_done = s->ps_done + su0;
    a.err = err; a.ps_rom = ctx->d_rom_ps; a.env_rom = ctx->d_rom_env; a.misc_rom = ctx->d_rom_misc; a.n_units = n; a.rot_nosat = ctx->ps_rot_nosat;
    LAUNCH("ps_frame_kernel", st, xb::launch_ps_frame(a, ctx->num_sms, st));
    y.ch_fac = 2;
    y.pcm_unit_stride = 4096;
    LAUNCH("qmf_synth_hq_kernel", st, xb::launch_qmf_synth_hq(y, ctx->num_sms, st));
    xb::QmfSynthArgs r = y;
    r.matrix = a.right; r.mat_stride = 4096; r.states = s->syn_states_r + u0 * 1280; r.pos = s->syn_pos_r + u0 * 2;
    r.params = a.synp_r; r.pcm = d_time_out + 1; r.gate = a.ps_done;
    LAUNCH("qmf_synth_hq_kernel", st, xb::launch_qmf_synth_hq(r, ctx->num_sms, st));
    ctx->launches += 3;
  } else {
    y.ch_fac = 1;
    LAUNCH("qmf_synth_hq_kernel", st, xb::launch_qmf_synth_hq(y, ctx->num_sms, st));
    ctx->launches += 1;
  }
  return XAAC_B200_OK;
}

static int32_t sbr_check(xaac_b200_ctx *ctx, xaac_b200_sbr_state *s) {
  if (!ctx || !s) return XAAC_B200_ERR_ARG;
  if (s->lp_only) return bad_arg(ctx, "state was created for the low-power stage (XAAC_B200_SBR_STATE_LP)");
  if (!ctx->have_qmf_rom || !ctx->have_env_rom || (s->with_ps && !ctx->have_ps_rom)) {
    snprintf(ctx->err, sizeof(ctx->err), "set_qmf_rom / set_env_rom / set_ps_rom have not all been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  return XAAC_B200_OK;
}

// whole batch of a state on the caller's stream, or (XAAC_B200_DEV_CHUNK) chunked over internal streams forked from / joined to it
static int32_t sbr_dec_all(xaac_b200_ctx *ctx, xaac_b200_sbr_state *s, const int16_t *d_side, const int16_t *d_time_in,
                           const int32_t *d_w32, const int8_t *d_adj, int16_t *d_time_out, int32_t *d_err, cudaStream_t stream) {
  const long long n = s->n_units;
  const long long chunk = ctx->dev_chunk;
  if (chunk <= 0 || n <= chunk) {
    int32_t rc = ensure_sbr_scratch(ctx, s, n);
    if (rc != XAAC_B200_OK) return rc;
    return sbr_dec_range(ctx, s, 0, n, d_side, d_time_in, d_time_out, d_err, stream, 0, d_w32, d_adj);
  }
  const int K = ctx->dev_streams;
  if (!ctx->dev_fork) {
    CK(cudaEventCreateWithFlags(&ctx->dev_fork, cudaEventDisableTiming), "cudaEventCreate");
    for (int k = 0; k < 8; k++) {
      CK(cudaStreamCreateWithFlags(&ctx->dev_st[k], cudaStreamNonBlocking), "cudaStreamCreate");
      CK(cudaEventCreateWithFlags(&ctx->dev_join[k], cudaEventDisableTiming), "cudaEventCreate");
    }
  }
  int32_t rc = ensure_sbr_scratch(ctx, s, (long long)K * chunk);
  if (rc != XAAC_B200_OK) return rc;
  CK(cudaEventRecord(ctx->dev_fork, stream), "cudaEventRecord");
  for (int k = 0; k < K; k++) CK(cudaStreamWaitEvent(ctx->dev_st[k], ctx->dev_fork, 0), "cudaStreamWaitEvent");
  const long long out_words = s->with_ps ? 4096 : 2048;
  int k = 0;
  for (long long u0 = 0; u0 < n; u0 += chunk, k = (k + 1) % K) {
    const long long m = n - u0 < chunk ? n - u0 : chunk;
    rc = sbr_dec_range(ctx, s, u0, m, d_side + u0 * xb::kSideWords, d_time_in ? d_time_in + u0 * 1024 : nullptr,
                       d_time_out + u0 * out_words, d_err ? d_err + u0 : nullptr, ctx->dev_st[k], (long long)k * chunk,
                       d_w32 ? d_w32 + u0 * 1024 : nullptr, d_adj ? d_adj + u0 : nullptr);
    if (rc != XAAC_B200_OK) break;  // the caller's stream still joins the chunks already queued
  }
  for (int q = 0; q < K; q++) {
    CK(cudaEventRecord(ctx->dev_join[q], ctx->dev_st[q]), "cudaEventRecord");
    CK(cudaStreamWaitEvent(stream, ctx->dev_join[q], 0), "cudaStreamWaitEvent");
  }
  return rc;
}

int32_t xaac_b200_sbr_dec_hq_dev(xaac_b200_ctx *ctx, xaac_b200_sbr_state *s, const int16_t *d_side,
                                 const int16_t *d_time_in, int16_t *d_time_out, int32_t *d_err, void *stream) {
  int32_t rc = sbr_check(ctx, s);
  if (rc != XAAC_B200_OK) return rc;
  if (!d_side || !d_time_in || !d_time_out) return bad_arg(ctx, "null buffer");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  return sbr_dec_all(ctx, s, d_side, d_time_in, nullptr, nullptr, d_time_out, d_err, (cudaStream_t)stream);
}

int32_t xaac_b200_sbr_dec_hq_w32_dev(xaac_b200_ctx *ctx, xaac_b200_sbr_state *s, const int16_t *d_side, const int32_t *d_w32,
                                     const int8_t *d_qshift_adj, int16_t *d_time_out, int32_t *d_err, void *stream) {
  int32_t rc = sbr_check(ctx, s);
  if (rc != XAAC_B200_OK) return rc;
  if (!d_side || !d_w32 || !d_qshift_adj || !d_time_out) return bad_arg(ctx, "null buffer");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  return sbr_dec_all(ctx, s, d_side, nullptr, d_w32, d_qshift_adj, d_time_out, d_err, (cudaStream_t)stream);
}

// One low-power frame for units [u0, u0 + n) of the state.
static int32_t sbr_dec_lp_range(xaac_b200_ctx *ctx, xaac_b200_sbr_state *s, long long u0, long long n, const int16_t *d_side,
                                const int16_t *d_time_in, int16_t *d_time_out, int32_t out_ch, int32_t *d_err,
                                cudaStream_t st, const int32_t *d_w32 = nullptr, const int8_t *d_adj = nullptr) {
  xb::SbrLpArgs a;
  a.w32 = d_w32; a.qshift_adj = d_adj;
  a.side = d_side; a.time_in = d_time_in; a.time_out = d_time_out;
  a.anal_states = s->anal_states + u0 * 320; a.anal_pos = s->anal_pos + u0 * 2; a.syn_pos = s->syn_pos + u0 * 2;
  a.sf = s->sf + u0 * 8; a.misc = s->misc + u0 * 16; a.env = s->env + u0 * xb::kEnvStWords;
  a.syn_states = s->syn_states + u0 * 1280; a.bw_prev = s->bw_prev + u0 * 6; a.lpc = s->lpc + u0 * 256; a.ov = s->ov + u0 * 768;
  a.err = d_err ? d_err : s->err + u0;
  a.lp_rom = ctx->d_rom_lp; a.env_rom = ctx->d_rom_env; a.misc_rom = ctx->d_rom_misc;
  a.n_units = n; a.out_ch = out_ch;
  LAUNCH("sbr_dec_lp_kernel", st, xb::launch_sbr_dec_lp(a, ctx->num_sms, st));
  ctx->launches += 1;
  return XAAC_B200_OK;
}

int32_t xaac_b200_sbr_dec_lp_dev(xaac_b200_ctx *ctx, xaac_b200_sbr_state *s, const int16_t *d_side,
                                 const int16_t *d_time_in, int16_t *d_time_out, int32_t out_ch, int32_t *d_err,
                                 void *stream) {
  if (!ctx || !s) return XAAC_B200_ERR_ARG;
  if (!ctx->have_qmf_rom || !ctx->have_env_rom) {
    snprintf(ctx->err, sizeof(ctx->err), "set_qmf_rom / set_env_rom have not both been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (!ctx->d_rom_lp) return bad_arg(ctx, "the installed QMF tables are not supported by the low-power kernel");
  if (!d_side || !d_time_in || !d_time_out) return bad_arg(ctx, "null buffer");
  if (out_ch < 1 || (s->n_units % out_ch) != 0) return bad_arg(ctx, "out_ch must divide the number of units");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  return sbr_dec_lp_range(ctx, s, 0, s->n_units, d_side, d_time_in, d_time_out, out_ch, d_err, (cudaStream_t)stream);
}

int32_t xaac_b200_sbr_dec_lp_w32_dev(xaac_b200_ctx *ctx, xaac_b200_sbr_state *s, const int16_t *d_side, const int32_t *d_w32,
                                     const int8_t *d_qshift_adj, int16_t *d_time_out, int32_t out_ch, int32_t *d_err,
                                     void *stream) {
  if (!ctx || !s) return XAAC_B200_ERR_ARG;
  if (!ctx->have_qmf_rom || !ctx->have_env_rom) {
    snprintf(ctx->err, sizeof(ctx->err), "set_qmf_rom / set_env_rom have not both been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (!ctx->d_rom_lp) return bad_arg(ctx, "the installed QMF tables are not supported by the low-power kernel");
  if (!d_side || !d_w32 || !d_qshift_adj || !d_time_out) return bad_arg(ctx, "null buffer");
  if (out_ch < 1 || (s->n_units % out_ch) != 0) return bad_arg(ctx, "out_ch must divide the number of units");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  return sbr_dec_lp_range(ctx, s, 0, s->n_units, d_side, nullptr, d_time_out, out_ch, d_err, (cudaStream_t)stream, d_w32,
                          d_qshift_adj);
}

// HE-AAC frame from host buffers: IMDCT (mono or one core channel per unit) -> WORD32->PCM16 hand-over -> SBR stage.
static int32_t xaac_b200_heaac_frame_host_impl(xaac_b200_ctx *ctx, xaac_b200_imdct_state *ist, xaac_b200_sbr_state *s,
                                   const int32_t *spec, const uint8_t *ics, const int16_t *side, int16_t *pcm,
                                   int32_t *err) {
  int32_t rc = sbr_check(ctx, s);
  if (rc != XAAC_B200_OK) return rc;
  if (!ist || !ctx->have_imdct_rom) return bad_arg(ctx, "IMDCT state / ROM missing");
  if (ist->n_units != s->n_units) return bad_arg(ctx, "IMDCT and SBR states must have the same number of units");
  if (!spec || !ics || !side || !pcm) return bad_arg(ctx, "null buffer");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  const int64_t n_units = s->n_units;
  int64_t chunk = ctx->host_chunk;
  if (chunk > n_units) chunk = n_units;
  const size_t out_words = s->with_ps ? 4096 : 2048;
  // per-unit staging: spec 4096 | WORD32 out 4096 | side 2464 | pcm16 in 2048 | pcm out 2*out_words | err 4 | ics 2 | adj 1 (+1)
  const size_t o_spec = 0, o_w32 = 4096, o_side = 8192, o_p16 = 8192 + 2464, o_pcm = o_p16 + 2048,
               o_err = o_pcm + 2 * out_words, o_ics = o_err + 4, o_adj = o_ics + 2, per_unit = o_adj + 2;
  rc = ensure_stage(ctx, per_unit * (size_t)chunk);
  if (rc != XAAC_B200_OK) return rc;
  // stage scratch for the chunks in flight only (one slot range per pipeline stream), not for the whole batch
  rc = ensure_sbr_scratch(ctx, s, (int64_t)xaac_b200_ctx::kPipe * chunk);
  if (rc != XAAC_B200_OK) return rc;
  int slot = 0;
  for (int64_t u0 = 0; u0 < n_units; u0 += chunk, slot = (slot + 1) % xaac_b200_ctx::kPipe) {
    const int64_t n = (n_units - u0 < chunk) ? (n_units - u0) : chunk;
    cudaStream_t st = ctx->streams[slot];
    uint8_t *base = (uint8_t *)ctx->stage[slot];
    int32_t *d_spec = (int32_t *)(base + o_spec * chunk), *d_w32 = (int32_t *)(base + o_w32 * chunk);
    int16_t *d_side = (int16_t *)(base + o_side * chunk);
    int16_t *d_pcm = (int16_t *)(base + o_pcm * chunk);
    int32_t *d_err = (int32_t *)(base + o_err * chunk);
    uint8_t *d_ics = base + o_ics * chunk;
    int8_t *d_adj = (int8_t *)(base + o_adj * chunk);
    CK(cudaMemcpyAsync(d_spec, spec + u0 * 1024, (size_t)n * 4096, cudaMemcpyHostToDevice, st), "H2D spec");
    CK(cudaMemcpyAsync(d_ics, ics + u0 * 2, (size_t)n * 2, cudaMemcpyHostToDevice, st), "H2D ics");
    CK(cudaMemcpyAsync(d_side, side + u0 * xb::kSideWords, (size_t)n * xb::kSideWords * 2, cudaMemcpyHostToDevice, st),
       "H2D side");
    rc = xaac_b200_imdct_process_dev(ctx, d_spec, ist->d_overlap + u0 * 512, ist->d_wstate + u0 * 2, d_ics, d_w32, d_adj, n,
                                     1, st);
    if (rc != XAAC_B200_OK) return rc;
    // the WORD32 -> PCM16 hand-over happens in the analysis bank's load
    rc = sbr_dec_range(ctx, s, u0, n, d_side, nullptr, d_pcm, d_err, st, (long long)slot * chunk, d_w32, d_adj);
    if (rc != XAAC_B200_OK) return rc;
    CK(cudaMemcpyAsync(pcm + u0 * out_words, d_pcm, (size_t)n * out_words * 2, cudaMemcpyDeviceToHost, st), "D2H pcm");
    if (err) CK(cudaMemcpyAsync(err + u0, d_err, (size_t)n * 4, cudaMemcpyDeviceToHost, st), "D2H err");
  }
  for (int i = 0; i < xaac_b200_ctx::kPipe; i++) CK(cudaStreamSynchronize(ctx->streams[i]), "stream sync");
  return XAAC_B200_OK;
}

// Stereo HE-AACv1 (low-power SBR) frames from host buffers: unit = one channel; IMDCT -> hand-over -> fused LP stage.
static int32_t xaac_b200_heaac_lp_frame_host_impl(xaac_b200_ctx *ctx, xaac_b200_imdct_state *ist, xaac_b200_sbr_state *s,
                                      const int32_t *spec, const uint8_t *ics, const int16_t *side, int16_t *pcm,
                                      int32_t out_ch, int32_t *err) {
  if (!ctx || !s) return XAAC_B200_ERR_ARG;
  if (!ctx->have_qmf_rom || !ctx->have_env_rom || !ctx->d_rom_lp) {
    snprintf(ctx->err, sizeof(ctx->err), "set_qmf_rom / set_env_rom have not both been called (or tables unsupported)");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (!ist || !ctx->have_imdct_rom) return bad_arg(ctx, "IMDCT state / ROM missing");
  if (ist->n_units != s->n_units) return bad_arg(ctx, "IMDCT and SBR states must have the same number of units");
  if (!spec || !ics || !side || !pcm) return bad_arg(ctx, "null buffer");
  if (out_ch < 1 || out_ch > 8 || (s->n_units % out_ch) != 0) return bad_arg(ctx, "out_ch must divide the number of units");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  const int64_t n_units = s->n_units;
  int64_t chunk = 4096;
  chunk -= chunk % out_ch;  // a chunk holds whole frames: the kernel derives frame / channel from the chunk-local index
  if (chunk > n_units) chunk = n_units;
  // per-unit staging: spec 4096 | WORD32 out 4096 | side 2464 | pcm16 in 2048 | pcm out 4096 | err 4 | ics 2 | adj 1 (+1)
  const size_t o_spec = 0, o_w32 = 4096, o_side = 8192, o_p16 = 8192 + 2464, o_pcm = o_p16 + 2048, o_err = o_pcm + 4096,
               o_ics = o_err + 4, o_adj = o_ics + 2, per_unit = o_adj + 2;
  int32_t rc = ensure_stage(ctx, per_unit * (size_t)chunk);
  if (rc != XAAC_B200_OK) return rc;
  int slot = 0;
  for (int64_t u0 = 0; u0 < n_units; u0 += chunk, slot = (slot + 1) % xaac_b200_ctx::kPipe) {
    const int64_t n = (n_units - u0 < chunk) ? (n_units - u0) : chunk;
    cudaStream_t st = ctx->streams[slot];
    uint8_t *base = (uint8_t *)ctx->stage[slot];
    int32_t *d_spec = (int32_t *)(base + o_spec * chunk), *d_w32 = (int32_t *)(base + o_w32 * chunk);
    int16_t *d_side = (int16_t *)(base + o_side * chunk);
    int16_t *d_pcm = (int16_t *)(base + o_pcm * chunk);
    int32_t *d_err = (int32_t *)(base + o_err * chunk);
    uint8_t *d_ics = base + o_ics * chunk;
    int8_t *d_adj = (int8_t *)(base + o_adj * chunk);
    CK(cudaMemcpyAsync(d_spec, spec + u0 * 1024, (size_t)n * 4096, cudaMemcpyHostToDevice, st), "H2D spec");
    CK(cudaMemcpyAsync(d_ics, ics + u0 * 2, (size_t)n * 2, cudaMemcpyHostToDevice, st), "H2D ics");
    CK(cudaMemcpyAsync(d_side, side + u0 * xb::kSideWords, (size_t)n * xb::kSideWords * 2, cudaMemcpyHostToDevice, st),
       "H2D side");
    rc = xaac_b200_imdct_process_dev(ctx, d_spec, ist->d_overlap + u0 * 512, ist->d_wstate + u0 * 2, d_ics, d_w32, d_adj, n,
                                     1, st);
    if (rc != XAAC_B200_OK) return rc;
    // the WORD32 -> PCM16 hand-over happens in the fused stage's load
    rc = sbr_dec_lp_range(ctx, s, u0, n, d_side, nullptr, d_pcm, out_ch, d_err, st, d_w32, d_adj);
    if (rc != XAAC_B200_OK) return rc;
    CK(cudaMemcpyAsync(pcm + u0 * 2048, d_pcm, (size_t)n * 4096, cudaMemcpyDeviceToHost, st), "D2H pcm");
    if (err) CK(cudaMemcpyAsync(err + u0, d_err, (size_t)n * 4, cudaMemcpyDeviceToHost, st), "D2H err");
  }
  for (int i = 0; i < xaac_b200_ctx::kPipe; i++) CK(cudaStreamSynchronize(ctx->streams[i]), "stream sync");
  return XAAC_B200_OK;
}

int32_t xaac_b200_imdct_out_to_pcm16_dev(xaac_b200_ctx *ctx, const int32_t *d_in, const int8_t *d_qshift_adj,
                                         int16_t *d_out, int64_t n_units, int32_t mode, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (n_units < 0 || (mode != 0 && mode != 1)) return bad_arg(ctx, "n_units/mode");
  if (n_units == 0) return XAAC_B200_OK;
  if (!d_in || !d_qshift_adj || !d_out) return bad_arg(ctx, "null buffer");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  LAUNCH("pcm16_from_imdct_kernel", stream, xb::launch_pcm16_from_imdct(d_in, d_qshift_adj, d_out, n_units, mode, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return XAAC_B200_OK;
}

int32_t xaac_b200_set_usac_rom(xaac_b200_ctx *ctx, const void *tables, size_t bytes) {
  if (!ctx || !tables) return bad_arg(ctx, "null");
  if (bytes < (size_t)xb::kURomBytes) return bad_arg(ctx, "USAC ROM blob shorter than 15880 bytes");
  if (xb::usac_fd_check_tables((const uint8_t *)tables) != 0) return bad_arg(ctx, "USAC tables");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  if (!ctx->d_rom_usac) CK(cudaMalloc((void **)&ctx->d_rom_usac, xb::kURomBytes + 56), "cudaMalloc(usac rom)");
  CK(cudaMemcpy(ctx->d_rom_usac, tables, xb::kURomBytes, cudaMemcpyHostToDevice), "H2D usac rom");
  return XAAC_B200_OK;
}

int32_t xaac_b200_usac_fd_frm_dec_dev(xaac_b200_ctx *ctx, const int32_t *d_coef, int32_t *d_overlap, uint8_t *d_wstate,
                                      const uint8_t *d_ics, int32_t *d_out, int64_t n_units, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!ctx->d_rom_usac) {
    snprintf(ctx->err, sizeof(ctx->err), "xaac_b200_set_usac_rom has not been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (n_units < 0) return bad_arg(ctx, "n_units");
  if (n_units == 0) return XAAC_B200_OK;
  if (!d_coef || !d_overlap || !d_wstate || !d_ics || !d_out) return bad_arg(ctx, "null buffer");
  xb::UsacFdArgs a;
  a.coef = d_coef; a.overlap = d_overlap; a.wstate = d_wstate; a.ics = d_ics; a.out = d_out; a.rom = ctx->d_rom_usac;
  a.n_units = n_units;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  LAUNCH("usac_fd_kernel", stream, xb::launch_usac_fd(a, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return XAAC_B200_OK;
}

int32_t xaac_b200_peak_limiter_state_init(int32_t *state, int32_t num_channels, int32_t sample_rate) {
  // ixheaacd_peak_limiter_init (decoder/ixheaacd_peak_limiter.c:45-75): host-side, no device involved
  if (!state || num_channels < 1 || num_channels > 2 || sample_rate < 1) return XAAC_B200_ERR_ARG;
  const uint32_t attack = (uint32_t)(5.0f * (float)sample_rate / 1000);
  if (attack < 1 || attack > (uint32_t)xb::kPlMaxAttack) return XAAC_B200_ERR_ARG;
  memset(state, 0, sizeof(int32_t) * xb::kPlWords);
  const float ac = (float)pow(0.1, 1.0 / (attack + 1));
  const float rc = (float)pow(0.1, 1.0 / (50.0f * sample_rate / 1000 + 1));
  const float one = 1.0f;
  const double done = 1.0;
  memcpy(state + xb::kPlAttackConst, &ac, 4);
  memcpy(state + xb::kPlReleaseConst, &rc, 4);
  memcpy(state + xb::kPlGainMod, &one, 4);
  memcpy(state + xb::kPlMinGain, &one, 4);
  memcpy(state + xb::kPlPsg, &done, 8);
  state[xb::kPlAttack] = (int32_t)attack;
  state[xb::kPlLimiterOn] = 1;
  state[xb::kPlNumCh] = num_channels;
  return XAAC_B200_OK;
}

static int32_t peak_limiter_launches(xaac_b200_ctx *ctx, const xb::PeakLimArgs &a, void *scratch, void *stream) {
  LAUNCH("peak_limiter_kernel", stream, xb::launch_peak_limiter(a, scratch, 0, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  if (scratch) {
    LAUNCH("peak_limiter_smooth_kernel", stream, xb::launch_peak_limiter(a, scratch, 1, ctx->num_sms, (cudaStream_t)stream));
    LAUNCH("peak_limiter_finish_kernel", stream, xb::launch_peak_limiter(a, scratch, 2, ctx->num_sms, (cudaStream_t)stream));
    ctx->launches += 2;
  }
  return XAAC_B200_OK;
}

int32_t xaac_b200_peak_limiter_dev(xaac_b200_ctx *ctx, int32_t *d_state, const int32_t *d_samples,
                                   const int8_t *d_qshift_adj, int32_t *d_out32, int16_t *d_pcm16, int32_t *d_err,
                                   int64_t n_units, int32_t num_channels, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (n_units < 0 || num_channels < 1 || num_channels > 2) return bad_arg(ctx, "n_units / num_channels");
  if (n_units == 0) return XAAC_B200_OK;
  if (!d_state || !d_samples || !d_qshift_adj || (!d_out32 && !d_pcm16)) return bad_arg(ctx, "null buffer");
  xb::PeakLimArgs a;
  a.state = d_state; a.samples = d_samples; a.qshift_adj = d_qshift_adj; a.out32 = d_out32; a.pcm16 = d_pcm16; a.err = d_err;
  a.n_units = n_units; a.ch = num_channels;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  // Streams whose attack / release recursion is active are finished by two follow-up kernels (recursion with lane = stream);
  // their raw gains travel through a stream-ordered scratch block (4 KB per stream).
  void *scratch = nullptr;
  if (n_units >= 64) {
    if (!ctx->pool) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = ctx->device;
      CK(cudaMemPoolCreate(&ctx->pool, &props), "cudaMemPoolCreate");
      unsigned long long keep = ~0ull;
      CK(cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep), "cudaMemPoolSetAttribute");
    }
    CK(cudaMallocFromPoolAsync(&scratch, xb::peak_limiter_scratch_bytes(n_units), ctx->pool, (cudaStream_t)stream),
       "cudaMallocFromPoolAsync");
  }
  const int32_t rc = peak_limiter_launches(ctx, a, scratch, stream);
  // the scratch block goes back to the pool on every path (stream-ordered: after the kernels that were queued)
  if (scratch && cudaFreeAsync(scratch, (cudaStream_t)stream) != cudaSuccess && rc == XAAC_B200_OK)
    return fail(ctx, cudaGetLastError(), "cudaFreeAsync");
  return rc;
}

int32_t xaac_b200_dec_sbrdata_dev(xaac_b200_ctx *ctx, int16_t *d_records, int64_t n_elements, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!ctx->have_env_rom || !ctx->d_rom_misc) {
    snprintf(ctx->err, sizeof(ctx->err), "xaac_b200_set_env_rom has not been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (n_elements < 0) return bad_arg(ctx, "n_elements");
  if (n_elements == 0) return XAAC_B200_OK;
  if (!d_records || ((uintptr_t)d_records & 15) != 0) return bad_arg(ctx, "records: null or not 16-byte aligned");
  LAUNCH("sbr_sideinfo_kernel", stream, xb::launch_sbr_sideinfo(d_records, n_elements, ctx->d_rom_misc, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return XAAC_B200_OK;
}

int32_t xaac_b200_decode_ps_data_dev(xaac_b200_ctx *ctx, int16_t *d_records, int64_t n_units, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (n_units < 0) return bad_arg(ctx, "n_units");
  if (n_units == 0) return XAAC_B200_OK;
  if (!d_records || ((uintptr_t)d_records & 15) != 0) return bad_arg(ctx, "records: null or not 16-byte aligned");
  LAUNCH("ps_sideinfo_kernel", stream, xb::launch_ps_sideinfo(d_records, n_units, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return XAAC_B200_OK;
}

int32_t xaac_b200_set_esbr_rom(xaac_b200_ctx *ctx, const void *tables, size_t bytes) {
  if (!ctx || !tables) return bad_arg(ctx, "null");
  if (bytes < (size_t)xb::kEsRomBytes) return bad_arg(ctx, "eSBR ROM blob shorter than 6288 bytes");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  const size_t n = xb::esbr_synth_table_bytes();
  uint8_t *img = (uint8_t *)calloc(1, n + 64);
  if (!img) return XAAC_B200_FATAL;
  ctx->esbr_periodic = xb::esbr_synth_build_tables((const uint8_t *)tables, img);
  cudaError_t e = ctx->d_rom_esbr ? cudaSuccess : cudaMalloc((void **)&ctx->d_rom_esbr, n);
  if (e == cudaSuccess) e = cudaMemcpy(ctx->d_rom_esbr, img, n, cudaMemcpyHostToDevice);
  free(img);
  if (e != cudaSuccess) return fail(ctx, e, "cudaMemcpy(esbr rom)");
  return XAAC_B200_OK;
}

static int32_t esbr_synth_common(xaac_b200_ctx *ctx, const float *d_qmf, int32_t *d_states, int32_t *d_pos, float *d_out,
                                 int16_t *d_pcm16, int32_t ch_fac, int32_t *d_err, int64_t n_units, void *stream);
int32_t xaac_b200_esbr_synth64_dev(xaac_b200_ctx *ctx, const float *d_qmf, int32_t *d_states, int32_t *d_pos, float *d_out,
                                   int32_t *d_err, int64_t n_units, void *stream) {
  if (ctx && n_units > 0 && !d_out) return bad_arg(ctx, "null buffer");
  return esbr_synth_common(ctx, d_qmf, d_states, d_pos, d_out, nullptr, 1, d_err, n_units, stream);
}
int32_t xaac_b200_esbr_synth64_pcm16_dev(xaac_b200_ctx *ctx, const float *d_qmf, int32_t *d_states, int32_t *d_pos,
                                         float *d_out, int16_t *d_pcm16, int32_t ch_fac, int32_t *d_err, int64_t n_units,
                                         void *stream) {
  if (ctx && n_units > 0 && !d_pcm16) return bad_arg(ctx, "null buffer");
  if (ctx && (ch_fac < 1 || ch_fac > 8 || n_units % ch_fac != 0)) return bad_arg(ctx, "ch_fac must be 1..8 and divide n_units");
  return esbr_synth_common(ctx, d_qmf, d_states, d_pos, d_out, d_pcm16, ch_fac, d_err, n_units, stream);
}
static int32_t esbr_synth_common(xaac_b200_ctx *ctx, const float *d_qmf, int32_t *d_states, int32_t *d_pos, float *d_out,
                                 int16_t *d_pcm16, int32_t ch_fac, int32_t *d_err, int64_t n_units, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!ctx->d_rom_esbr) {
    snprintf(ctx->err, sizeof(ctx->err), "xaac_b200_set_esbr_rom has not been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (n_units < 0) return bad_arg(ctx, "n_units");
  if (n_units == 0) return XAAC_B200_OK;
  if (!d_qmf || !d_states || !d_pos || (!d_out && !d_pcm16)) return bad_arg(ctx, "null buffer");
  xb::EsbrSynthArgs a;
  a.qmf = d_qmf; a.states = d_states; a.pos = d_pos; a.out = d_out; a.err = d_err; a.rom = ctx->d_rom_esbr;
  a.pcm16 = d_pcm16; a.pcm_ch_fac = ch_fac;
  a.n_units = n_units; a.periodic = ctx->esbr_periodic;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  LAUNCH("esbr_synth_kernel", stream, xb::launch_esbr_synth(a, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return XAAC_B200_OK;
}

static int32_t esbr_anal_common(xaac_b200_ctx *ctx, const float *d_time_in, const int32_t *d_core, const int16_t *d_pcm,
                                int32_t ch_fac, int32_t *d_states, int32_t *d_pos, float *d_qmf, int32_t *d_err,
                                int64_t n_units, void *stream);
int32_t xaac_b200_esbr_anal32_dev(xaac_b200_ctx *ctx, const float *d_time_in, int32_t *d_states, int32_t *d_pos, float *d_qmf,
                                  int32_t *d_err, int64_t n_units, void *stream) {
  if (ctx && n_units > 0 && !d_time_in) return bad_arg(ctx, "null buffer");
  return esbr_anal_common(ctx, d_time_in, nullptr, nullptr, 1, d_states, d_pos, d_qmf, d_err, n_units, stream);
}
int32_t xaac_b200_esbr_anal32_core_dev(xaac_b200_ctx *ctx, const int32_t *d_core, int32_t *d_states, int32_t *d_pos,
                                       float *d_qmf, int32_t *d_err, int64_t n_units, void *stream) {
  if (ctx && n_units > 0 && !d_core) return bad_arg(ctx, "null buffer");
  return esbr_anal_common(ctx, nullptr, d_core, nullptr, 1, d_states, d_pos, d_qmf, d_err, n_units, stream);
}
int32_t xaac_b200_esbr_anal32_pcm16_dev(xaac_b200_ctx *ctx, const int16_t *d_pcm16, int32_t ch_fac, int32_t *d_states,
                                        int32_t *d_pos, float *d_qmf, int32_t *d_err, int64_t n_units, void *stream) {
  if (ctx && n_units > 0 && !d_pcm16) return bad_arg(ctx, "null buffer");
  if (ctx && (ch_fac < 1 || ch_fac > 8 || n_units % ch_fac != 0)) return bad_arg(ctx, "ch_fac must be 1..8 and divide n_units");
  return esbr_anal_common(ctx, nullptr, nullptr, d_pcm16, ch_fac, d_states, d_pos, d_qmf, d_err, n_units, stream);
}
static int32_t esbr_anal_common(xaac_b200_ctx *ctx, const float *d_time_in, const int32_t *d_core, const int16_t *d_pcm,
                                int32_t ch_fac, int32_t *d_states, int32_t *d_pos, float *d_qmf, int32_t *d_err,
                                int64_t n_units, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!ctx->d_rom_esbr) {
    snprintf(ctx->err, sizeof(ctx->err), "xaac_b200_set_esbr_rom has not been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (n_units < 0) return bad_arg(ctx, "n_units");
  if (n_units == 0) return XAAC_B200_OK;
  if ((!d_time_in && !d_core && !d_pcm) || !d_states || !d_pos || !d_qmf) return bad_arg(ctx, "null buffer");
  xb::EsbrAnalArgs a;
  a.core_in = d_core; a.pcm_in = d_pcm; a.pcm_ch_fac = ch_fac;
  a.time_in = d_time_in; a.states = d_states; a.pos = d_pos; a.qmf = d_qmf; a.err = d_err; a.rom = ctx->d_rom_esbr;
  a.n_units = n_units; a.periodic = ctx->esbr_periodic;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  LAUNCH("esbr_anal_kernel", stream, xb::launch_esbr_anal(a, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return XAAC_B200_OK;
}

int32_t xaac_b200_esbr_generate_hf_dev(xaac_b200_ctx *ctx, const float *d_src_re, const float *d_src_im, const float *d_pv_re,
                                       const float *d_pv_im, float *d_dst_re, float *d_dst_im, const int32_t *d_par,
                                       float *d_bw_prev, int32_t *d_patch_out, int32_t *d_err, int64_t n_units, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (n_units < 0) return bad_arg(ctx, "n_units");
  if (n_units == 0) return XAAC_B200_OK;
  if (!d_src_re || !d_src_im || !d_dst_re || !d_dst_im || !d_par || !d_bw_prev) return bad_arg(ctx, "null buffer");
  if ((d_pv_re == nullptr) != (d_pv_im == nullptr)) return bad_arg(ctx, "pv_re / pv_im must both be given or both be null");
  xb::EsbrHfgenArgs a;
  a.src_re = d_src_re; a.src_im = d_src_im; a.pv_re = d_pv_re; a.pv_im = d_pv_im; a.dst_re = d_dst_re; a.dst_im = d_dst_im;
  a.par = d_par; a.bw_prev = d_bw_prev; a.patch_out = d_patch_out; a.err = d_err; a.n_units = n_units;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  LAUNCH("esbr_hfgen_kernel", stream, xb::launch_esbr_hfgen(a, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return XAAC_B200_OK;
}

int32_t xaac_b200_set_esbr_envcalc_rom(xaac_b200_ctx *ctx, const void *random_phase, size_t bytes) {
  if (!ctx || !random_phase) return bad_arg(ctx, "null");
  if (bytes < (size_t)xb::kEecRphaseBytes) return bad_arg(ctx, "ixheaac_random_phase blob shorter than 4096 bytes");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  cudaError_t e = ctx->d_rom_rphase ? cudaSuccess : cudaMalloc((void **)&ctx->d_rom_rphase, xb::kEecRphaseBytes);
  if (e == cudaSuccess) e = cudaMemcpy(ctx->d_rom_rphase, random_phase, xb::kEecRphaseBytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return fail(ctx, e, "cudaMemcpy(esbr random phase)");
  return XAAC_B200_OK;
}

int32_t xaac_b200_esbr_env_calc_dev(xaac_b200_ctx *ctx, float *d_re, float *d_im, int32_t *d_ipar, const float *d_fpar,
                                    float *d_state, int32_t *d_err, int64_t n_units, void *stream) {
  return xaac_b200_esbr_env_calc_tes_dev(ctx, d_re, d_im, nullptr, nullptr, 40, d_ipar, d_fpar, d_state, d_err, n_units, stream);
}

int32_t xaac_b200_esbr_env_calc_tes_dev(xaac_b200_ctx *ctx, float *d_re, float *d_im, const float *d_low_re, const float *d_low_im,
                                        int32_t low_rows, int32_t *d_ipar, const float *d_fpar, float *d_state, int32_t *d_err,
                                        int64_t n_units, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if ((d_low_re == nullptr) != (d_low_im == nullptr) || (low_rows != 40 && low_rows != 72)) return bad_arg(ctx, "low-band arrays");
  if (!ctx->d_rom_rphase) {
    snprintf(ctx->err, sizeof(ctx->err), "xaac_b200_set_esbr_envcalc_rom has not been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (n_units < 0) return bad_arg(ctx, "n_units");
  if (n_units == 0) return XAAC_B200_OK;
  if (!d_re || !d_im || !d_ipar || !d_fpar || !d_state) return bad_arg(ctx, "null buffer");
  xb::EsbrEnvcalcArgs a;
  a.re = d_re; a.im = d_im; a.ipar = d_ipar; a.fpar = d_fpar; a.state = d_state; a.rphase = ctx->d_rom_rphase; a.err = d_err;
  a.n_units = n_units; a.low_re = d_low_re; a.low_im = d_low_im; a.low_stride = (long long)low_rows * 64;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  LAUNCH("esbr_envcalc_kernel", stream, xb::launch_esbr_envcalc(a, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return XAAC_B200_OK;
}

int32_t xaac_b200_set_hbe_rom(xaac_b200_ctx *ctx, const void *tables, size_t bytes) {
  if (!ctx || !tables) return bad_arg(ctx, "null");
  if (bytes < (size_t)xb::kHromWords * 4) return bad_arg(ctx, "harmonic-transposer ROM blob shorter than 37296 bytes");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  cudaError_t e = ctx->d_rom_hbe ? cudaSuccess : cudaMalloc((void **)&ctx->d_rom_hbe, (size_t)xb::kHromWords * 4);
  if (e == cudaSuccess) e = cudaMemcpy(ctx->d_rom_hbe, tables, (size_t)xb::kHromWords * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return fail(ctx, e, "cudaMemcpy(hbe rom)");
  return XAAC_B200_OK;
}

static int32_t hbe_launch(xaac_b200_ctx *ctx, const float *d_qmf_re, const float *d_qmf_im, int64_t in_stride, float *d_pv_re,
                          float *d_pv_im, int64_t out_stride, const int32_t *d_cfg, float *d_state, int32_t *d_err,
                          int64_t n_units, void *stream) {
  xb::EsbrHbeArgs a;
  a.qmf_re = d_qmf_re; a.qmf_im = d_qmf_im; a.pv_re = d_pv_re; a.pv_im = d_pv_im; a.in_stride = in_stride; a.out_stride = out_stride;
  a.cfg = d_cfg; a.state = d_state; a.err = d_err; a.rom = ctx->d_rom_hbe; a.n_units = n_units;
  LAUNCH("esbr_hbe_kernel", stream, xb::launch_esbr_hbe(a, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return XAAC_B200_OK;
}

int32_t xaac_b200_esbr_hbe_apply_dev(xaac_b200_ctx *ctx, const float *d_qmf_re, const float *d_qmf_im, float *d_pv_re,
                                     float *d_pv_im, const int32_t *d_cfg, float *d_state, int32_t *d_err, int64_t n_units,
                                     void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!ctx->d_rom_hbe) {
    snprintf(ctx->err, sizeof(ctx->err), "xaac_b200_set_hbe_rom has not been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (n_units < 0) return bad_arg(ctx, "n_units");
  if (n_units == 0) return XAAC_B200_OK;
  if (!d_qmf_re || !d_qmf_im || !d_pv_re || !d_pv_im || !d_cfg || !d_state) return bad_arg(ctx, "null buffer");
  return hbe_launch(ctx, d_qmf_re, d_qmf_im, 2048, d_pv_re, d_pv_im, 2048, d_cfg, d_state, d_err, n_units, stream);
}

int32_t xaac_b200_set_fps_rom(xaac_b200_ctx *ctx, const void *tables, size_t bytes) {
  if (!ctx || !tables) return bad_arg(ctx, "null");
  if (bytes < (size_t)xb::kFpsRomWords * 4) return bad_arg(ctx, "float-PS ROM blob shorter than 4064 bytes");
  {
    const int32_t *w = (const int32_t *)tables;  // the kernel unrolls the serial all-pass links with these lengths
    if (w[xb::kFpsRomDser] != 3 || w[xb::kFpsRomDser + 1] != 4 || w[xb::kFpsRomDser + 2] != 5)
      return bad_arg(ctx, "float-PS ROM: delay_sample_ser is not {3, 4, 5}");
    for (int i = 0; i < 64; i++)
      if (w[xb::kFpsRomQdelN + i] < 1 || w[xb::kFpsRomQdelN + i] > 14) return bad_arg(ctx, "float-PS ROM: qmf_delay_idx_tbl out of 1..14");
  }
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  cudaError_t e = ctx->d_rom_fps ? cudaSuccess : cudaMalloc((void **)&ctx->d_rom_fps, (size_t)xb::kFpsRomWords * 4);
  if (e == cudaSuccess) e = cudaMemcpy(ctx->d_rom_fps, tables, (size_t)xb::kFpsRomWords * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return fail(ctx, e, "cudaMemcpy(fps rom)");
  return XAAC_B200_OK;
}

int32_t xaac_b200_esbr_ps_apply_dev(xaac_b200_ctx *ctx, const float *d_low_re, const float *d_low_im, int32_t low_rows,
                                    const float *d_high_re, const float *d_high_im, const int32_t *d_rg_par,
                                    const float *d_side, float *d_state, float *d_left, float *d_right, int32_t *d_err,
                                    int64_t n_units, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!ctx->d_rom_fps) {
    snprintf(ctx->err, sizeof(ctx->err), "xaac_b200_set_fps_rom has not been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (n_units < 0) return bad_arg(ctx, "n_units");
  if (n_units == 0) return XAAC_B200_OK;
  if (low_rows != 40 && low_rows != 72) return bad_arg(ctx, "low_rows must be 40 or 72");
  if (!d_low_re || !d_low_im || !d_high_re || !d_high_im || !d_rg_par || !d_side || !d_state || !d_left || !d_right)
    return bad_arg(ctx, "null buffer");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  xb::EsbrPsArgs a;
  a.low_re = d_low_re; a.low_im = d_low_im; a.high_re = d_high_re; a.high_im = d_high_im; a.rg_par = d_rg_par;
  a.low_stride = (long long)low_rows * 64; a.side = d_side; a.state = d_state; a.left = d_left; a.right = d_right; a.err = d_err;
  a.rom = ctx->d_rom_fps; a.n_units = n_units;
  LAUNCH("esbr_ps_kernel", stream, xb::launch_esbr_ps(a, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return XAAC_B200_OK;
}

// Whole eSBR stage (eSBR branch of ixheaacd_sbr_dec for USAC mono / stereo channels without harmonic transposer, PS, MPS):
// analysis bank (+ history shift, + core hand-over) -> HF generator (+ history shift) -> envelope adjuster -> synthesis bank
// (+ regrouping, + optional PCM16 hand-over).  Four launches on one stream.
static int32_t esbr_dec_impl(xaac_b200_ctx *ctx, const xaac_b200_esbr_state_view *st, float *pv_re, float *pv_im,
                             float *hbe_state, const int32_t *d_hbe_cfg, const float *d_time_in, const int32_t *d_core_in,
                             const int32_t *d_hf_par, int32_t *d_ec_ipar, const float *d_ec_fpar, const int32_t *d_rg_par,
                             float *d_out, int16_t *d_pcm16, int32_t ch_fac, int32_t *d_err, int64_t n_units, void *stream,
                             const xaac_b200_esbr_ps_view *ps = nullptr, const float *d_ps_side = nullptr,
                             float *d_out_r = nullptr, int phases = 3) {
  // phases: 1 = front (analysis bank, transposer, HF generator), 2 = back (envelope adjuster, PS, synthesis), 3 = both
  const bool front = (phases & 1) != 0, back = (phases & 2) != 0;
  const bool hbe = pv_re != nullptr;
  if (!ctx->d_rom_esbr || !ctx->d_rom_rphase || (hbe && !ctx->d_rom_hbe) || (ps && !ctx->d_rom_fps)) {
    snprintf(ctx->err, sizeof(ctx->err),
             "xaac_b200_set_esbr_rom / _set_esbr_envcalc_rom / _set_hbe_rom / _set_fps_rom have not been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (n_units < 0) return bad_arg(ctx, "n_units");
  if (n_units == 0) return XAAC_B200_OK;
  if (!st || !st->qmf_re || !st->qmf_im || !st->out_re || !st->out_im || !st->anal_states || !st->anal_pos ||
      !st->synth_states || !st->synth_pos || !st->bw_prev || !st->patch || !st->ec_state)
    return bad_arg(ctx, "state view with null members");
  if (front && ((!d_time_in && !d_core_in) || !d_hf_par)) return bad_arg(ctx, "null buffer");
  if (back && (!d_ec_ipar || !d_ec_fpar || !d_rg_par || (!d_out && !d_pcm16))) return bad_arg(ctx, "null buffer");
  if (back && d_pcm16 && (ch_fac < 1 || ch_fac > 8 || n_units % ch_fac != 0)) return bad_arg(ctx, "ch_fac must be 1..8 and divide n_units");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  cudaStream_t s = (cudaStream_t)stream;
  const long long low_stride = hbe ? 72 * 64 : 40 * 64;
  if (front) {
    xb::EsbrAnalArgs a;
    a.time_in = d_time_in; a.core_in = d_time_in ? nullptr : d_core_in; a.states = st->anal_states; a.pos = st->anal_pos;
    a.qmf = nullptr; a.stage_re = st->qmf_re; a.stage_im = st->qmf_im; a.err = d_err; a.rom = ctx->d_rom_esbr;
    a.n_units = n_units; a.periodic = ctx->esbr_periodic; a.stage_hist_rows = hbe ? 40 : 8;
    LAUNCH("esbr_anal_kernel", stream, xb::launch_esbr_anal(a, ctx->num_sms, s));
  }
  if (front && hbe) {  // sbr_dec.c:896-907: the frame's new slots (rows 40..71) -> ph_vocod_qmf rows 8..39
    xb::EsbrHbeArgs a;
    a.qmf_re = st->qmf_re + 40 * 64; a.qmf_im = st->qmf_im + 40 * 64; a.in_stride = low_stride;
    a.pv_re = pv_re + 8 * 64; a.pv_im = pv_im + 8 * 64; a.out_stride = 40 * 64;
    a.cfg = d_hbe_cfg; a.state = hbe_state; a.err = d_err ? d_err + 4 * n_units : nullptr; a.rom = ctx->d_rom_hbe;
    a.n_units = n_units; a.shift_rows = 1;
    LAUNCH("esbr_hbe_kernel", stream, xb::launch_esbr_hbe(a, ctx->num_sms, s));
    ctx->launches++;
  }
  if (front) {
    xb::EsbrHfgenArgs a;
    a.src_re = st->qmf_re; a.src_im = st->qmf_im; a.pv_re = pv_re; a.pv_im = pv_im; a.dst_re = st->out_re; a.dst_im = st->out_im;
    a.par = d_hf_par; a.bw_prev = st->bw_prev; a.patch_out = st->patch; a.err = d_err ? d_err + n_units : nullptr;
    a.n_units = n_units; a.shift_rows = 1; a.src_stride = low_stride;
    LAUNCH("esbr_hfgen_kernel", stream, xb::launch_esbr_hfgen(a, ctx->num_sms, s));
    ctx->launches += 2;
  }
  if (!back) return XAAC_B200_OK;
  {
    xb::EsbrEnvcalcArgs a;
    a.re = st->out_re; a.im = st->out_im; a.ipar = d_ec_ipar; a.fpar = d_ec_fpar; a.state = st->ec_state;
    a.rphase = ctx->d_rom_rphase; a.err = d_err ? d_err + 2 * n_units : nullptr; a.n_units = n_units;
    a.low_re = st->qmf_re; a.low_im = st->qmf_im; a.low_stride = low_stride;  // inter-TES reads the low band
    LAUNCH("esbr_envcalc_kernel", stream, xb::launch_esbr_envcalc(a, ctx->num_sms, s));
  }
  if (ps) {  // sbr_dec.c:976-1001: regrouping + PS, then both channels' synthesis banks
    if (!ps->ps_state || !ps->left || !ps->right || !ps->synth_states_r || !ps->synth_pos_r || !d_ps_side || !d_out || !d_out_r)
      return bad_arg(ctx, "PS view with null members");
    xb::EsbrPsArgs a;
    a.low_re = st->qmf_re; a.low_im = st->qmf_im; a.high_re = st->out_re; a.high_im = st->out_im; a.rg_par = d_rg_par;
    a.low_stride = low_stride; a.side = d_ps_side; a.state = ps->ps_state; a.left = ps->left; a.right = ps->right;
    a.err = d_err ? d_err + 5 * n_units : nullptr; a.rom = ctx->d_rom_fps; a.n_units = n_units;
    LAUNCH("esbr_ps_kernel", stream, xb::launch_esbr_ps(a, ctx->num_sms, s));
    for (int ch = 0; ch < 2; ch++) {
      xb::EsbrSynthArgs y;
      y.qmf = ch ? ps->right : ps->left; y.states = ch ? ps->synth_states_r : st->synth_states;
      y.pos = ch ? ps->synth_pos_r : st->synth_pos; y.out = ch ? d_out_r : d_out;
      y.err = (d_err && !ch) ? d_err + 3 * n_units : nullptr;
      y.rom = ctx->d_rom_esbr; y.n_units = n_units; y.periodic = ctx->esbr_periodic;
      LAUNCH("esbr_synth_kernel", stream, xb::launch_esbr_synth(y, ctx->num_sms, s));
    }
    ctx->launches += 4;
    return XAAC_B200_OK;
  }
  {
    xb::EsbrSynthArgs a;
    a.qmf = nullptr; a.states = st->synth_states; a.pos = st->synth_pos; a.out = d_out; a.err = d_err ? d_err + 3 * n_units : nullptr;
    a.rom = ctx->d_rom_esbr; a.n_units = n_units; a.periodic = ctx->esbr_periodic;
    a.rg_low_re = st->qmf_re; a.rg_low_im = st->qmf_im; a.rg_high_re = st->out_re; a.rg_high_im = st->out_im; a.rg_par = d_rg_par;
    a.rg_low_stride = low_stride;
    a.pcm16 = d_pcm16; a.pcm_ch_fac = d_pcm16 ? ch_fac : 1;
    LAUNCH("esbr_synth_kernel", stream, xb::launch_esbr_synth(a, ctx->num_sms, s));
  }
  ctx->launches += 2;
  return XAAC_B200_OK;
}

int32_t xaac_b200_esbr_dec_dev(xaac_b200_ctx *ctx, const xaac_b200_esbr_state_view *st, const float *d_time_in,
                               const int32_t *d_core_in, const int32_t *d_hf_par, int32_t *d_ec_ipar, const float *d_ec_fpar,
                               const int32_t *d_rg_par, float *d_out, int16_t *d_pcm16, int32_t ch_fac, int32_t *d_err,
                               int64_t n_units, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  return esbr_dec_impl(ctx, st, nullptr, nullptr, nullptr, nullptr, d_time_in, d_core_in, d_hf_par, d_ec_ipar, d_ec_fpar, d_rg_par,
                       d_out, d_pcm16, ch_fac, d_err, n_units, stream);
}

// The same stage with the harmonic transposer (hbe_flag = 1): five launches.
int32_t xaac_b200_esbr_dec_hbe_dev(xaac_b200_ctx *ctx, const xaac_b200_esbr_hbe_state_view *st, const float *d_time_in,
                                   const int32_t *d_core_in, const int32_t *d_hbe_cfg, const int32_t *d_hf_par,
                                   int32_t *d_ec_ipar, const float *d_ec_fpar, const int32_t *d_rg_par, float *d_out,
                                   int16_t *d_pcm16, int32_t ch_fac, int32_t *d_err, int64_t n_units, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!st || !st->pv_re || !st->pv_im || !st->hbe_state || !d_hbe_cfg) return bad_arg(ctx, "harmonic-transposer state / cfg missing");
  return esbr_dec_impl(ctx, &st->base, st->pv_re, st->pv_im, st->hbe_state, d_hbe_cfg, d_time_in, d_core_in, d_hf_par, d_ec_ipar,
                       d_ec_fpar, d_rg_par, d_out, d_pcm16, ch_fac, d_err, n_units, stream);
}

// Mono + PS element: the stage above up to the envelope adjuster, the PS kernel, two synthesis banks.
int32_t xaac_b200_esbr_dec_ps_dev(xaac_b200_ctx *ctx, const xaac_b200_esbr_hbe_state_view *st, const xaac_b200_esbr_ps_view *ps,
                                  const float *d_time_in, const int32_t *d_core_in, const int32_t *d_hbe_cfg,
                                  const int32_t *d_hf_par, int32_t *d_ec_ipar, const float *d_ec_fpar, const int32_t *d_rg_par,
                                  const float *d_ps_side, float *d_out_l, float *d_out_r, int32_t *d_err, int64_t n_units,
                                  void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!st || !ps) return bad_arg(ctx, "state views missing");
  const bool hbe = st->pv_re != nullptr;
  if (hbe && (!st->pv_im || !st->hbe_state || !d_hbe_cfg)) return bad_arg(ctx, "harmonic-transposer state / cfg missing");
  return esbr_dec_impl(ctx, &st->base, hbe ? st->pv_re : nullptr, hbe ? st->pv_im : nullptr, hbe ? st->hbe_state : nullptr,
                       d_hbe_cfg, d_time_in, d_core_in, d_hf_par, d_ec_ipar, d_ec_fpar, d_rg_par, d_out_l, nullptr, 1, d_err,
                       n_units, stream, ps, d_ps_side, d_out_r);
}

// The stage in two halves, for frames on which the host has to look at the HF generator's patch table before the envelope
// adjuster runs (reset frames and frames where sbr_patching_mode changes: ixheaacd_createlimiterbands, esbr_envcal.c:169-190).
int32_t xaac_b200_esbr_dec_front_dev(xaac_b200_ctx *ctx, const xaac_b200_esbr_hbe_state_view *st, const float *d_time_in,
                                     const int32_t *d_core_in, const int32_t *d_hbe_cfg, const int32_t *d_hf_par, int32_t *d_err,
                                     int64_t n_units, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!st) return bad_arg(ctx, "state view missing");
  const bool hbe = st->pv_re != nullptr;
  if (hbe && (!st->pv_im || !st->hbe_state || !d_hbe_cfg)) return bad_arg(ctx, "harmonic-transposer state / cfg missing");
  return esbr_dec_impl(ctx, &st->base, hbe ? st->pv_re : nullptr, hbe ? st->pv_im : nullptr, hbe ? st->hbe_state : nullptr,
                       d_hbe_cfg, d_time_in, d_core_in, d_hf_par, nullptr, nullptr, nullptr, nullptr, nullptr, 1, d_err, n_units,
                       stream, nullptr, nullptr, nullptr, 1);
}

int32_t xaac_b200_esbr_dec_back_dev(xaac_b200_ctx *ctx, const xaac_b200_esbr_hbe_state_view *st, const xaac_b200_esbr_ps_view *ps,
                                    int32_t *d_ec_ipar, const float *d_ec_fpar, const int32_t *d_rg_par, const float *d_ps_side,
                                    float *d_out, float *d_out_r, int16_t *d_pcm16, int32_t ch_fac, int32_t *d_err,
                                    int64_t n_units, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!st) return bad_arg(ctx, "state view missing");
  if (ps && d_pcm16) return bad_arg(ctx, "the mono + PS stage has float outputs only");
  const bool hbe = st->pv_re != nullptr;
  return esbr_dec_impl(ctx, &st->base, hbe ? st->pv_re : nullptr, hbe ? st->pv_im : nullptr, hbe ? st->hbe_state : nullptr, nullptr,
                       nullptr, nullptr, nullptr, d_ec_ipar, d_ec_fpar, d_rg_par, d_out, d_pcm16, ch_fac, d_err, n_units, stream, ps,
                       d_ps_side, d_out_r, 2);
}

// apply_processing = 0 (the first frames of a stream, before the SBR header has been seen): the stage only upsamples — history
// shifts, analysis bank, sbr_qmf_out cleared, synthesis bank(s) over the regrouped core bands (sbr_dec.c:836-878, 964-1003).
int32_t xaac_b200_esbr_dec_bypass_dev(xaac_b200_ctx *ctx, const xaac_b200_esbr_hbe_state_view *st, const xaac_b200_esbr_ps_view *ps,
                                      const float *d_time_in, const int32_t *d_core_in, const int32_t *d_rg_par, float *d_out,
                                      float *d_out_r, int16_t *d_pcm16, int32_t ch_fac, int32_t *d_err, int64_t n_units,
                                      void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!ctx->d_rom_esbr) {
    snprintf(ctx->err, sizeof(ctx->err), "xaac_b200_set_esbr_rom has not been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (n_units < 0) return bad_arg(ctx, "n_units");
  if (n_units == 0) return XAAC_B200_OK;
  if (!st || !st->base.qmf_re || !st->base.qmf_im || !st->base.out_re || !st->base.out_im || !st->base.anal_states ||
      !st->base.anal_pos || !st->base.synth_states || !st->base.synth_pos)
    return bad_arg(ctx, "state view with null members");
  if ((!d_time_in && !d_core_in) || !d_rg_par || (!d_out && !d_pcm16)) return bad_arg(ctx, "null buffer");
  if (ps && (!ps->synth_states_r || !ps->synth_pos_r || !d_out_r || !d_out || d_pcm16)) return bad_arg(ctx, "PS view / outputs");
  if (d_pcm16 && (ch_fac < 1 || ch_fac > 8 || n_units % ch_fac != 0)) return bad_arg(ctx, "ch_fac must be 1..8 and divide n_units");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  cudaStream_t s = (cudaStream_t)stream;
  const bool hbe = st->pv_re != nullptr;
  const long long low_stride = hbe ? 72 * 64 : 40 * 64;
  {
    xb::EsbrAnalArgs a;
    a.time_in = d_time_in; a.core_in = d_time_in ? nullptr : d_core_in; a.states = st->base.anal_states; a.pos = st->base.anal_pos;
    a.qmf = nullptr; a.stage_re = st->base.qmf_re; a.stage_im = st->base.qmf_im; a.err = d_err; a.rom = ctx->d_rom_esbr;
    a.n_units = n_units; a.periodic = ctx->esbr_periodic; a.stage_hist_rows = hbe ? 40 : 8;
    LAUNCH("esbr_anal_kernel", stream, xb::launch_esbr_anal(a, ctx->num_sms, s));
  }
  if (hbe) {  // sbr_dec.c:858-867: the phase-vocoder arrays still move their last 8 rows to the front
    if (!st->pv_im) return bad_arg(ctx, "pv_im");
    CK(cudaMemcpy2DAsync(st->pv_re, 2560 * 4, st->pv_re + 32 * 64, 2560 * 4, 8 * 64 * 4, (size_t)n_units, cudaMemcpyDeviceToDevice, s), "shift");
    CK(cudaMemcpy2DAsync(st->pv_im, 2560 * 4, st->pv_im + 32 * 64, 2560 * 4, 8 * 64 * 4, (size_t)n_units, cudaMemcpyDeviceToDevice, s), "shift");
  }
  // sbr_dec.c:964-969 clears all of sbr_qmf_out (the shift of its history rows before that is moot)
  CK(cudaMemsetAsync(st->base.out_re, 0, (size_t)n_units * 2560 * 4, s), "memset");
  CK(cudaMemsetAsync(st->base.out_im, 0, (size_t)n_units * 2560 * 4, s), "memset");
  for (int ch = 0; ch < (ps ? 2 : 1); ch++) {
    xb::EsbrSynthArgs a;
    a.qmf = nullptr; a.states = ch ? ps->synth_states_r : st->base.synth_states; a.pos = ch ? ps->synth_pos_r : st->base.synth_pos;
    a.out = ch ? d_out_r : d_out; a.err = (d_err && !ch) ? d_err + 3 * n_units : nullptr;
    a.rom = ctx->d_rom_esbr; a.n_units = n_units; a.periodic = ctx->esbr_periodic;
    a.rg_low_re = st->base.qmf_re; a.rg_low_im = st->base.qmf_im; a.rg_high_re = st->base.out_re; a.rg_high_im = st->base.out_im;
    a.rg_par = d_rg_par; a.rg_low_stride = low_stride;
    a.pcm16 = ch ? nullptr : d_pcm16; a.pcm_ch_fac = d_pcm16 ? ch_fac : 1;
    LAUNCH("esbr_synth_kernel", stream, xb::launch_esbr_synth(a, ctx->num_sms, s));
  }
  ctx->launches += ps ? 3 : 2;
  return XAAC_B200_OK;
}

int32_t xaac_b200_set_block_rom(xaac_b200_ctx *ctx, const void *block_tables, size_t bytes) {
  if (!ctx || !block_tables) return bad_arg(ctx, "null");
  if (bytes < (size_t)xb::kBromBytes) return bad_arg(ctx, "block-tables ROM shorter than 620 bytes");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  cudaError_t e = ctx->d_rom_block ? cudaSuccess : cudaMalloc((void **)&ctx->d_rom_block, 640);
  if (e == cudaSuccess) e = cudaMemcpy(ctx->d_rom_block, block_tables, xb::kBromBytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return fail(ctx, e, "cudaMemcpy(block rom)");
  return XAAC_B200_OK;
}

int32_t xaac_b200_aac_spectral_dev(xaac_b200_ctx *ctx, int32_t *d_spec, const uint8_t *d_side, int32_t *d_pns_seed, int32_t *d_err,
                                   int64_t n_units, void *stream) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  if (!ctx->d_rom_block) {
    snprintf(ctx->err, sizeof(ctx->err), "xaac_b200_set_block_rom has not been called");
    return XAAC_B200_ERR_NO_ROM;
  }
  if (n_units < 0) return bad_arg(ctx, "n_units");
  if (n_units == 0) return XAAC_B200_OK;
  if (!d_spec || !d_side || !d_err) return bad_arg(ctx, "null buffer (d_err carries the per-element verdict between the two kernels)");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  xb::AacSpectralArgs a;
  a.spec = d_spec; a.side = d_side; a.err = d_err; a.pns_seed = d_pns_seed; a.rom = ctx->d_rom_block; a.n_units = n_units;
  LAUNCH("aac_spectral_kernels", stream, xb::launch_aac_spectral(a, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches += d_pns_seed ? 3 : 2;
  return XAAC_B200_OK;
}

int32_t xaac_b200_kernel_timing(xaac_b200_ctx *ctx, int32_t enable) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  CK(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  for (int i = 0; i < ctx->n_ticks; i++) {
    cudaEventDestroy(ctx->tick_ev[i][0]);
    cudaEventDestroy(ctx->tick_ev[i][1]);
  }
  ctx->n_ticks = 0;
  ctx->timing = enable != 0;
  return XAAC_B200_OK;
}

int32_t xaac_b200_kernel_times(xaac_b200_ctx *ctx, char *buf, size_t buf_bytes) {
  if (!ctx || !buf || buf_bytes < 2) return XAAC_B200_ERR_ARG;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  CK(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  const char *names[64];
  double ms[64];
  int cnt[64], n = 0;
  for (int i = 0; i < ctx->n_ticks; i++) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, ctx->tick_ev[i][0], ctx->tick_ev[i][1]) != cudaSuccess) continue;
    int k = 0;
    while (k < n && strcmp(names[k], ctx->tick_name[i]) != 0) k++;
    if (k == n) {
      if (n == 64) continue;
      names[n] = ctx->tick_name[i]; ms[n] = 0; cnt[n] = 0; n++;
    }
    ms[k] += t;
    cnt[k]++;
  }
  size_t o = 0;
  buf[0] = 0;
  for (int k = 0; k < n; k++) {
    int w = snprintf(buf + o, buf_bytes - o, "%s:%.6f:%d;", names[k], ms[k], cnt[k]);
    if (w < 0 || (size_t)w >= buf_bytes - o) break;
    o += (size_t)w;
  }
  return XAAC_B200_OK;
}

int32_t xaac_b200_dev_alloc(xaac_b200_ctx *ctx, size_t bytes, void **d_ptr) {
  if (!ctx || !d_ptr) return bad_arg(ctx, "dev_alloc");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  CK(cudaMalloc(d_ptr, bytes ? bytes : 1), "cudaMalloc");
  return XAAC_B200_OK;
}
int32_t xaac_b200_dev_free(xaac_b200_ctx *ctx, void *d_ptr) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  if (d_ptr) CK(cudaFree(d_ptr), "cudaFree");
  return XAAC_B200_OK;
}
int32_t xaac_b200_h2d(xaac_b200_ctx *ctx, void *d_dst, const void *h_src, size_t bytes) {
  if (!ctx || !d_dst || !h_src) return bad_arg(ctx, "h2d");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  CK(cudaMemcpy(d_dst, h_src, bytes, cudaMemcpyHostToDevice), "cudaMemcpy H2D");
  return XAAC_B200_OK;
}
int32_t xaac_b200_d2h(xaac_b200_ctx *ctx, void *h_dst, const void *d_src, size_t bytes) {
  if (!ctx || !h_dst || !d_src) return bad_arg(ctx, "d2h");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  CK(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost), "cudaMemcpy D2H");
  return XAAC_B200_OK;
}
int32_t xaac_b200_dev_memset(xaac_b200_ctx *ctx, void *d_ptr, int32_t value, size_t bytes) {
  if (!ctx || !d_ptr) return bad_arg(ctx, "dev_memset");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  CK(cudaMemset(d_ptr, value, bytes), "cudaMemset");
  CK(cudaStreamSynchronize(0), "sync");
  return XAAC_B200_OK;
}

int32_t xaac_b200_ipc_export(xaac_b200_ctx *ctx, void *d_ptr, void *handle64) {
  if (!ctx || !d_ptr || !handle64) return bad_arg(ctx, "ipc_export");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, d_ptr), "cudaIpcGetMemHandle");
  memcpy(handle64, &h, 64);
  return XAAC_B200_OK;
}
int32_t xaac_b200_ipc_import(xaac_b200_ctx *ctx, const void *handle64, void **d_ptr) {
  if (!ctx || !d_ptr || !handle64) return bad_arg(ctx, "ipc_import");
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");  // the mapping (and the lazily enabled peer access) is for THIS device
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
  return XAAC_B200_OK;
}
int32_t xaac_b200_ipc_close(xaac_b200_ctx *ctx, void *d_ptr) {
  if (!ctx) return XAAC_B200_ERR_ARG;
  CK(cudaSetDevice(ctx->device), "cudaSetDevice");
  if (d_ptr) CK(cudaIpcCloseMemHandle(d_ptr), "cudaIpcCloseMemHandle");
  return XAAC_B200_OK;
}

// *_host entry points: on any error the pipeline streams are drained before returning
int32_t xaac_b200_imdct_process_host(xaac_b200_ctx *ctx, xaac_b200_imdct_state *state, const int32_t *spec,
                                     const uint8_t *ics, int32_t *out, int8_t *qshift_adj, int32_t ch_fac) {
  return drain(ctx, xaac_b200_imdct_process_host_impl(ctx, state, spec, ics, out, qshift_adj, ch_fac));
}

int32_t xaac_b200_qmf_synth_hq_host(xaac_b200_ctx *ctx, xaac_b200_qmf_synth_state *state, const int32_t *matrix,
                                    const int16_t *params, int16_t *pcm, int32_t ch_fac) {
  return drain(ctx, xaac_b200_qmf_synth_hq_host_impl(ctx, state, matrix, params, pcm, ch_fac));
}

int32_t xaac_b200_heaac_frame_host(xaac_b200_ctx *ctx, xaac_b200_imdct_state *ist, xaac_b200_sbr_state *s,
                                   const int32_t *spec, const uint8_t *ics, const int16_t *side, int16_t *pcm,
                                   int32_t *err) {
  return drain(ctx, xaac_b200_heaac_frame_host_impl(ctx, ist, s, spec, ics, side, pcm, err));
}

int32_t xaac_b200_heaac_lp_frame_host(xaac_b200_ctx *ctx, xaac_b200_imdct_state *ist, xaac_b200_sbr_state *s,
                                      const int32_t *spec, const uint8_t *ics, const int16_t *side, int16_t *pcm,
                                      int32_t out_ch, int32_t *err) {
  return drain(ctx, xaac_b200_heaac_lp_frame_host_impl(ctx, ist, s, spec, ics, side, pcm, out_ch, err));
}

}  // extern "C"
