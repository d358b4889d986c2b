// synth_sim.cu — CPU lane simulator of qmf_synth_hq_kernel (TEST INFRASTRUCTURE, never linked into libxaac_b200.so).
//
// The kernel's arithmetic and index maps live in libxaac_b200/csrc/qmf_synth_core.cuh as __host__ __device__ functions
// written per lane.  This file runs them on the host, lane after lane and phase after phase, over a host copy of the
// warp's shared-memory rows, so that tests/test_synth_sim.py can compare the restructured data flow (lane = slot
// modulation in registers, linear-time window, ring <-> row mapping) with the oracle without a GPU.  What it cannot
// cover is the data movement (bulk copies, barriers) — the -m gpu tests do.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../libxaac_b200/csrc/qmf_synth_core.cuh"

using namespace xb::syn;

namespace {
constexpr int kQRomW32 = 0, kQRomSinCosL64 = 192, kQRomAltSinL64 = 320, kQRomQmfC = 904;
void build_tw(const uint8_t *qrom, SynTw *t) {
  const int16_t *w32 = reinterpret_cast<const int16_t *>(qrom + kQRomW32);
  const int16_t *sc = reinterpret_cast<const int16_t *>(qrom + kQRomSinCosL64);
  const int16_t *al = reinterpret_cast<const int16_t *>(qrom + kQRomAltSinL64);
  auto hi = [](int16_t v) { return (int32_t)((uint32_t)(uint16_t)v << 16); };
  for (int n = 0; n < 32; n++) t->pre[n] = make_int2(hi(sc[2 * n]), hi(sc[2 * n + 1]));
  for (int n = 0; n < 16; n++) t->alt[n] = make_int2(hi(al[2 * n]), hi(al[2 * n + 1]));
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 3; j++) t->w1[3 * i + j] = make_int2(hi(w32[6 * i + 2 * j]), hi(w32[6 * i + 2 * j + 1]));
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 3; j++)
      t->w2[3 * i + j] = make_int2(hi(w32[48 + 6 * i + 2 * j]), hi(w32[48 + 6 * i + 2 * j + 1]));
}
}  // namespace

// one unit; params = {ov_lb_scale, lb_scale, hb_scale, st_syn_scale, lsb, usb, split, 0}; returns 1 if the exact path ran
extern "C" int synth_sim_unit(const uint8_t *qrom, const int32_t *matrix, int16_t *fs, int16_t *pos,
                              const int16_t *prm, int16_t *pcm, int fast_bits, int ch_fac) {
  SynTw tw;
  build_tw(qrom, &tw);
  alignas(16) static int32_t rows[kRows * kRowW];
  alignas(16) int2 sh[2][64];
  int4 *rows4 = reinterpret_cast<int4 *>(rows);
  memset(rows, 0x5a, sizeof(rows));
  const int ov_lb_scale = prm[0], lb_scale = prm[1], hb_scale = prm[2], st_syn = prm[3];
  const int lsb = prm[4], usb = prm[5], split = prm[6];
  const int off0 = pos[0], fpos0 = pos[1];
  const int Bw0 = off0 >> 7, fp0 = fpos0 >> 6;
  const int ov_lb_shift = (st_syn - ov_lb_scale) - 8, lb_shift = (st_syn - lb_scale) - 8;
  const int hb_shift = (st_syn - hb_scale) - 8;
  const FoldK fk = fold_consts(-(st_syn - 3) + 1);
  int32_t *st32 = reinterpret_cast<int32_t *>(fs);
  const int32_t *c32 = reinterpret_cast<const int32_t *>(qrom + kQRomQmfC);
  WinCoef wc[32];
  for (int lane = 0; lane < 32; lane++) {
    for (int B = 0; B < 10; B++) history_store(rows4, lane, B, Bw0, st32[32 * (2 * B) + lane], st32[32 * (2 * B + 1) + lane]);
    for (int k = 0; k < 2; k++) {
      const int band = lane + 32 * k;
      sh[0][band] = shift_entry(band < lsb ? ov_lb_shift : (band < usb ? hb_shift : 0));
      sh[1][band] = shift_entry(band < lsb ? lb_shift : (band < usb ? hb_shift : 0));
    }
    window_coefs(wc[lane], c32, lane, Bw0, fp0);
    memcpy(rows + (kHist + lane) * kRowW, matrix + 128 * lane, 512);  // the bulk copy
  }
  bool bad = false;
  for (int lane = 0; lane < 32; lane++) {
    int32_t mx = 0, mn = 0;
    shift_row(rows + (kHist + lane) * kRowW, sh[lane < split ? 0 : 1], mx, mn);
    const int32_t lim = (int32_t)(1u << fast_bits);
    bad |= (mx >= lim) || (mn < -lim);
  }
  for (int lane = 0; lane < 32; lane++) {
    if (bad)
      slot_modulate<true>(rows + (kHist + lane) * kRowW, tw, fk, 0);
    else
      slot_modulate<false>(rows + (kHist + lane) * kRowW, tw, fk, 0);
  }
  for (int lane = 0; lane < 32; lane++) {
    if (Bw0 & 1)
      window_unit<1>(rows4, lane, wc[lane], pcm, ch_fac);
    else
      window_unit<0>(rows4, lane, wc[lane], pcm, ch_fac);
  }
  for (int lane = 0; lane < 32; lane++) {
    int B = Bw0 + 8;
    if (B >= 10) B -= 10;
    for (int r = 22; r < 32; r++) {
      int32_t w0, w1;
      state_words(rows4, lane, r, w0, w1);
      st32[32 * (2 * B) + lane] = w0;
      st32[32 * (2 * B + 1) + lane] = w1;
      B = B ? B - 1 : 9;
    }
  }
  int off = off0 + 1024, fpos = fpos0 + 128;
  if (off >= 1280) off -= 1280;
  if (fpos >= 640) fpos -= 640;
  pos[0] = (int16_t)off;
  pos[1] = (int16_t)fpos;
  return bad ? 1 : 0;
}
