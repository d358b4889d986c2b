// Float parametric stereo of the eSBR branch (ixheaacd_esbr_apply_ps, decoder/ixheaacd_ps_dec_flt.c:381-505): one warp owns one
// mono+PS frame (32 QMF slots x 64 bands) and carries it through
//   regrouping (ixheaacd_esbr_synthesis_regrp, sbr_dec.c:297-397) + the six look-ahead slots (sbr_dec.c:485-506)
//   hybrid analysis of bands 0..2 (ixheaacd_hyb_anal / ixheaacd_k_chan_filt, ps_dec_flt.c:77-201), 20-band configuration
//   transient detection + all-pass decorrelation (ixheaacd_esbr_ps_de_correlate, ps_dec_flt.c:507-865)
//   rotation with the per-envelope mixing matrices (ixheaacd_esbr_ps_apply_rotation, ps_dec_flt.c:867-1224)
//   hybrid synthesis (ixheaacd_hyb_synth, ps_dec_flt.c:203-228)
// and leaves the left / right QMF matrices in the layout the float synthesis bank kernel reads.  The mixing matrices h11..h22
// come from the host (they need the C library's double-precision cos / sin / atan2, ps_dec_flt.c:920-1010).
//
// Every sum keeps the reference's association; the file is compiled with -fmad=false so no product is contracted into an FMA.
// Lane maps: time slot (hybrid analysis / synthesis, band powers), parameter bin (transient detector), sub-subband or QMF band
// (decorrelator + rotation, serial over the slots because each is a recursion in time).
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace xb {
namespace {

constexpr int kPsWarps = 4;
constexpr int kWWords = 3 * 44 * 2;       // hybrid filter input, bands 0..2
constexpr int kHyStride = 13;             // 12 sub-subbands per slot, odd stride
constexpr int kHyWords = 4 * 32 * kHyStride;
constexpr int kPowStride = 21;
constexpr int kPowWords = 32 * kPowStride;
constexpr int kPStride = 65;
constexpr int kPWords = 32 * kPStride;
constexpr int kRingWords = 28 * 32;       // one lane's delay line + serial all-pass states: 28 floats, [slot][lane]
constexpr int kHsWords = 6 * 160;          // the frame's mixing matrices: previous + up to 5 envelopes, [set][8][20]
constexpr int kPsWarpWords = kWWords + kHyWords + kPowWords + kPWords + kRingWords + kHsWords;
// ring slots of one lane, all-pass (sub)bands: delay line re 0..1, im 2..3; link m of length {3, 4, 5}: re at 4 + {0, 3, 7},
// im at 16 + {0, 3, 7}.  Plain-delay bands (QMF bands >= 23): re 0..13, im 14..27.

struct Left {
  const float *lre, *lim, *hre, *him;
  int xo_first, xo_rest, stop;
  // regrouped cell (slot s < 32, band k): sbr_dec.c:297-397 for stereo_config_idx <= 0
  __device__ __forceinline__ void at(int s, int k, float &re, float &im) const {
    const int xo = s < stop ? xo_first : xo_rest;
    const int o = (2 + s) * 64 + k;
    const float *pr = k < xo ? lre : hre, *pi = k < xo ? lim : him;  // a select, not a branch: the loads of a batch overlap
    re = __ldg(pr + o); im = __ldg(pi + o);
  }
  // slots 32..37 (bands 0..4) always come from the low-band array (sbr_dec.c:485-506)
  __device__ __forceinline__ void ahead(int s, int k, float &re, float &im) const {
    const int o = (2 + s) * 64 + k;
    re = lre[o]; im = lim[o];
  }
  __device__ __forceinline__ void any(int s, int k, float &re, float &im) const {
    if (s < 32) at(s, k, re, im); else ahead(s, k, re, im);
  }
};

struct Mix {  // one bin's interpolated matrix
  float r11, r12, r21, r22, i11, i12, i21, i22;
  float d11r, d12r, d21r, d22r, d11i, d12i, d21i, d22i;
  // ps_dec_flt.c:1030-1075: start from the previous envelope's matrix, step = (target - start) / L
  __device__ __forceinline__ void start(const float *hs, int env, int bin, bool neg, int L) {
    const float *a = hs + env * 160 + bin;                      // previous (set env), target (set env + 1)
    const float *b = a + 160;
    const float sg = neg ? -1.0f : 1.0f;
    r11 = a[0]; r12 = a[20]; r21 = a[40]; r22 = a[60];
    i11 = sg * a[80]; i12 = sg * a[100]; i21 = sg * a[120]; i22 = sg * a[140];
    const float fl = (float)L;
    d11r = (b[0] - r11) / fl; d12r = (b[20] - r12) / fl; d21r = (b[40] - r21) / fl; d22r = (b[60] - r22) / fl;
    d11i = (sg * b[80] - i11) / fl; d12i = (sg * b[100] - i12) / fl; d21i = (sg * b[120] - i21) / fl; d22i = (sg * b[140] - i22) / fl;
  }
  __device__ __forceinline__ void step() {
    r11 += d11r; r12 += d12r; r21 += d21r; r22 += d22r;
    i11 += d11i; i12 += d12i; i21 += d21i; i22 += d22i;
  }
  __device__ __forceinline__ void apply(float &lr, float &li, float &rr, float &ri) const {
    const float a = r11 * lr - i11 * li + r21 * rr - i21 * ri;
    const float b = i11 * lr + r11 * li + i21 * rr + r21 * ri;
    const float c = r12 * lr - i12 * li + r22 * rr - i22 * ri;
    const float d = i12 * lr + r12 * li + i22 * rr + r22 * ri;
    lr = a; li = b; rr = c; ri = d;
  }
};

}  // namespace

__global__ void __launch_bounds__(kPsWarps * 32, 2) esbr_ps_kernel(const EsbrPsArgs p) {
  extern __shared__ float smem[];
  float *rom = smem;  // kFpsRomWords
  for (int i = threadIdx.x; i < kFpsRomWords; i += blockDim.x) rom[i] = p.rom[i];
  __syncthreads();
  const int *irom = reinterpret_cast<const int *>(rom);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *W = smem + kFpsRomWords + warp * kPsWarpWords;
  float *HL_re = W + kWWords, *HL_im = HL_re + 32 * kHyStride, *HR_re = HL_im + 32 * kHyStride, *HR_im = HR_re + 32 * kHyStride;
  float *POW = HR_im + 32 * kHyStride;
  float *PT = POW + kPowWords;
  float *RING = PT + kPWords + lane;
  float *HS = PT + kPWords + kRingWords;
  const int *grb = irom + kFpsRomGrb, *bgm = irom + kFpsRomBgm, *qdeln = irom + kFpsRomQdelN;

  for (long long u = (long long)blockIdx.x * kPsWarps + warp; u < p.n_units; u += (long long)gridDim.x * kPsWarps) {
    const float *side = p.side + u * kFpsSideWords;
    const int *iside = reinterpret_cast<const int *>(side);
    float *st = p.state + u * kFpsStWords;
    int *ist = reinterpret_cast<int *>(st) + kFpsStIdx;
    const int num_env = iside[kFpsSideNumEnv], usb = iside[kFpsSideUsb];
    int bad = (num_env < 1 || num_env > 5 || usb < 0 || usb > 64) ? 1 : 0;
    if (!bad) {
      if (iside[kFpsSideBorder] != 0 || iside[kFpsSideBorder + num_env] != 32) bad = 1;
      for (int e = 0; e < num_env; e++)
        if (iside[kFpsSideBorder + e + 1] <= iside[kFpsSideBorder + e]) bad = 1;
    }
    if (bad) {  // outside the kernel's subset (the reference would leave stale cells in the right matrix): refuse, touch nothing
      if (p.err && lane == 0) p.err[u] = -2;
      continue;
    }
    if (p.err && lane == 0) p.err[u] = 0;
    Left L;
    L.lre = p.low_re + u * p.low_stride; L.lim = p.low_im + u * p.low_stride;
    L.hre = p.high_re + u * 2560; L.him = p.high_im + u * 2560;
    L.xo_first = p.rg_par[4 * u]; L.xo_rest = p.rg_par[4 * u + 1]; L.stop = p.rg_par[4 * u + 2];
    float *out_l = p.left + u * 4096, *out_r = p.right + u * 4096;

    {  // the mixing matrices of the frame -> shared memory (every band lane of a bin interpolates them again)
      const int cnt = (num_env + 1) * 160;
#pragma unroll 6
      for (int i = lane; i < cnt; i += 32) HS[i] = __ldg(side + kFpsSideH + i);
    }
    // ps_dec_flt.c:419-431: bands above usb forget their decorrelator history
    for (int sb = lane; sb < 64; sb += 32)
      if (sb >= usb) {
        for (int m = 0; m < 3; m++)
          for (int k = 0; k < 3 + m; k++) {
            st[kFpsStSerQ + (m * 5 + k) * 64 + sb] = 0.f;
            st[kFpsStSerQ + 960 + (m * 5 + k) * 64 + sb] = 0.f;
          }
        for (int k = 0; k < 14; k++) {
          st[kFpsStQDel + k * 64 + sb] = 0.f;
          st[kFpsStQDel + 896 + k * 64 + sb] = 0.f;
        }
      }

    // ---- hybrid filter input: 12 slots of history + rows 6..37 of bands 0..2 (ps_dec_flt.c:148-165)
    for (int b = 0; b < 3; b++) {
      if (lane < 12) {
        W[b * 88 + lane] = st[kFpsStHyb + b * 12 + lane];
        W[b * 88 + 44 + lane] = st[kFpsStHyb + 60 + b * 12 + lane];
      }
      float re, im;
      L.any(lane + 6, b, re, im);
      W[b * 88 + 12 + lane] = re;
      W[b * 88 + 44 + 12 + lane] = im;
    }
    __syncwarp();
    // new history of bands 0..4 (the 20- and the 34-band hybrid keep the same samples, ps_dec_flt.c:432-437): rows 26..37
    if (lane < 12)
      for (int b = 0; b < 5; b++) {
        float re, im;
        L.any(26 + lane, b, re, im);
        st[kFpsStHyb + b * 12 + lane] = re;
        st[kFpsStHyb + 60 + b * 12 + lane] = im;
      }
    // ---- hybrid analysis, lane = slot
    {
      const float *p8 = rom + kFpsRomP8, *p2 = rom + kFpsRomP2, *cs8 = rom + kFpsRomCs8, *c2 = rom + kFpsRomCos2;
      float xr[13], xi[13];
      for (int n = 0; n < 13; n++) { xr[n] = W[lane + n]; xi[n] = W[44 + lane + n]; }
      for (int q = 0; q < 8; q++) {
        float re = 0.f, im = 0.f;
        for (int n = 0; n < 13; n++) {
          const float cv = cs8[q * 26 + 2 * n], sv = cs8[q * 26 + 2 * n + 1];
          re += p8[n] * (xr[n] * cv - xi[n] * sv);
          im += p8[n] * (xi[n] * cv + xr[n] * sv);
        }
        HL_re[lane * kHyStride + q] = re;
        HL_im[lane * kHyStride + q] = im;
      }
      for (int b = 1; b < 3; b++) {
        for (int n = 0; n < 13; n++) { xr[n] = W[b * 88 + lane + n]; xi[n] = W[b * 88 + 44 + lane + n]; }
        for (int q = 0; q < 2; q++) {
          float re = 0.f, im = 0.f;
          for (int n = 0; n < 13; n++) {
            const float cv = c2[q * 13 + n];
            re += p2[n] * (xr[n] * cv);
            im += p2[n] * (xi[n] * cv);
          }
          HL_re[lane * kHyStride + 6 + 2 * b + q] = re;
          HL_im[lane * kHyStride + 6 + 2 * b + q] = im;
        }
      }
      // ps_dec_flt.c:439-452
      float *r = HL_re + lane * kHyStride, *i = HL_im + lane * kHyStride;
      r[3] += r[4]; i[3] += i[4]; r[4] = 0.f; i[4] = 0.f;
      r[2] += r[5]; i[2] += i[5]; r[5] = 0.f; i[5] = 0.f;
    }
    // ---- |x|^2 of the QMF bands 3..63, lane = band
    for (int k0 = 0; k0 < 32; k0 += 4) {
      float re[8], im[8];
#pragma unroll
      for (int j = 0; j < 8; j++) L.at(k0 + (j >> 1), lane + 32 * (j & 1), re[j], im[j]);
#pragma unroll
      for (int j = 0; j < 8; j++) PT[(k0 + (j >> 1)) * kPStride + lane + 32 * (j & 1)] = re[j] * re[j] + im[j] * im[j];
    }
    __syncwarp();
    // ---- band powers per parameter bin, lane = slot (ps_dec_flt.c:610-636)
    {
      float *pw = POW + lane * kPowStride;
      for (int b = 0; b < 20; b++) pw[b] = 0.f;
      for (int gr = 0; gr < 10; gr++) {
        const int bin = bgm[gr] & 0xfff, sb = grb[gr];
        const float re = HL_re[lane * kHyStride + sb], im = HL_im[lane * kHyStride + sb];
        pw[bin] += re * re + im * im;
      }
      for (int gr = 10; gr < 22; gr++) {
        const int bin = bgm[gr] & 0xfff;
        float acc = pw[bin];
        for (int sb = grb[gr]; sb < grb[gr + 1]; sb++) acc += PT[lane * kPStride + sb];
        pw[bin] = acc;
      }
    }
    __syncwarp();
    // ---- transient detector, lane = bin; the ratio replaces the power (ps_dec_flt.c:638-663)
    if (lane < 20) {
      float peak = st[kFpsStBins + lane], nrg = st[kFpsStBins + 20 + lane], diff = st[kFpsStBins + 40 + lane];
      for (int k = 0; k < 32; k++) {
        const float pw = POW[k * kPowStride + lane];
        peak *= 0.765928338364649f;
        if (peak < pw) peak = pw;
        diff += 0.25f * (peak - pw - diff);
        nrg += 0.25f * (pw - nrg);
        const float qd = 1.5f * diff;
        POW[k * kPowStride + lane] = (qd <= nrg) ? 1.0f : nrg / qd;
      }
      st[kFpsStBins + lane] = peak; st[kFpsStBins + 20 + lane] = nrg; st[kFpsStBins + 40 + lane] = diff;
    }
    __syncwarp();
    const int l_delay0 = ist[0];
    const int ser0[3] = {ist[1], ist[2], ist[3]};
    int l_delay_end = l_delay0, ser_end[3] = {ser0[0], ser0[1], ser0[2]};
    constexpr int d0 = 3, d1 = 4, d2 = 5;  // delay_sample_ser (checked against the ROM by xaac_b200_set_fps_rom)
    // ---- decorrelator (ps_dec_flt.c:655-712 / 736-803) fused with the rotation (ps_dec_flt.c:1077-1100 / 1150-1186), one
    // (sub)band per lane, serial over the slots: pass 0 = the 10 sub-subband groups, passes 1, 2 = QMF bands 3..34 and 35..63.
    // The lane's delay line and all-pass states live in shared memory for the pass (28 floats), the pass's 32 input bands too.
    for (int pass = 0; pass < 3; pass++) {
      int sb, gr;
      bool active;
      if (pass == 0) {
        gr = lane; active = lane < 10; sb = active ? grb[gr] : 0;
      } else {
        sb = (pass == 1 ? 3 : 35) + lane; active = sb < 64;
        gr = 10;
        if (active) while (sb >= grb[gr + 1]) gr++;
        __syncwarp();
        const int sbl = active ? sb : 63;
        for (int k0 = 0; k0 < 32; k0 += 8) {  // input tile [slot][re 32 | im 32], coalesced rows, 8 rows in flight
          float re[8], im[8];
#pragma unroll
          for (int j = 0; j < 8; j++) L.at(k0 + j, sbl, re[j], im[j]);
#pragma unroll
          for (int j = 0; j < 8; j++) { PT[(k0 + j) * 64 + lane] = re[j]; PT[(k0 + j) * 64 + 32 + lane] = im[j]; }
        }
      }
      const bool plain = pass != 0 && sb >= 23;
      const int ds = pass == 0 ? 12 : 64;
      float *g_d = st + (pass == 0 ? kFpsStSubDel : kFpsStQDel) + sb, *g_s = st + (pass == 0 ? kFpsStSerSub : kFpsStSerQ) + sb;
      const int d_im = pass == 0 ? 24 : 896, s_im = pass == 0 ? 180 : 960;
      const int qnum = plain ? qdeln[sb] : 1;
      int qidx = plain ? ist[4 + sb] : 0;
      float fr = 0.f, fi = 0.f, sr[3] = {0.f, 0.f, 0.f}, si[3] = {0.f, 0.f, 0.f}, c[3] = {0.f, 0.f, 0.f};
      if (active) {
        {  // all 28 words of the lane's state in flight at once (link lengths 3, 4, 5: checked by xaac_b200_set_fps_rom)
          float v[28];
          if (plain) {
#pragma unroll
            for (int r = 0; r < 14; r++) {
              v[r] = r < qnum ? g_d[r * ds] : 0.f;
              v[14 + r] = r < qnum ? g_d[d_im + r * ds] : 0.f;
            }
          } else {
#pragma unroll
            for (int r = 0; r < 2; r++) { v[r] = g_d[r * ds]; v[2 + r] = g_d[d_im + r * ds]; }
#pragma unroll
            for (int m = 0; m < 3; m++)
#pragma unroll
              for (int r = 0; r < 3 + m; r++) {
                v[4 + (m == 0 ? 0 : m == 1 ? 3 : 7) + r] = g_s[(m * 5 + r) * ds];
                v[16 + (m == 0 ? 0 : m == 1 ? 3 : 7) + r] = g_s[s_im + (m * 5 + r) * ds];
              }
          }
#pragma unroll
          for (int r = 0; r < 28; r++) RING[32 * r] = v[r];
        }
        if (!plain) {
          float dsf = 1.0f;
          if (pass == 0) {
            fr = rom[kFpsRomSubRe + sb]; fi = rom[kFpsRomSubIm + sb];
          } else {
            fr = rom[kFpsRomQfRe + sb]; fi = rom[kFpsRomQfIm + sb];
            dsf = (sb <= 3) ? 1.0f : 1.0f + 3.0f * 0.05f - 0.05f * (float)sb;  // ps_dec_flt.c:738-744, decay_cutoff = 3
            dsf = dsf > 0.0f ? dsf : 0.0f;
          }
          for (int m = 0; m < 3; m++) {
            sr[m] = rom[(pass == 0 ? kFpsRomSSerRe : kFpsRomQSerRe) + sb * 3 + m];
            si[m] = rom[(pass == 0 ? kFpsRomSSerIm : kFpsRomQSerIm) + sb * 3 + m];
            c[m] = dsf * rom[kFpsRomDecay + m];
          }
        }
        const int bin = bgm[gr] & 0xfff;
        const bool neg = (bgm[gr] & 0x1000) != 0;
        int ld = l_delay0, sd0 = ser0[0], sd1 = ser0[1], sd2 = ser0[2];
        Mix H;
        for (int env = 0; env < num_env; env++) {
          const int k0 = iside[kFpsSideBorder + env], k1 = iside[kFpsSideBorder + env + 1];
          H.start(HS, env, bin, neg, k1 - k0);
          for (int k = k0; k < k1; k++) {
            float lr, li;
            if (pass == 0) { lr = HL_re[k * kHyStride + sb]; li = HL_im[k * kHyStride + sb]; }
            else { lr = PT[k * 64 + lane]; li = PT[k * 64 + 32 + lane]; }
            float r0, i0;
            if (plain) {
              r0 = RING[32 * qidx]; i0 = RING[32 * (14 + qidx)];
              RING[32 * qidx] = lr; RING[32 * (14 + qidx)] = li;
              if (++qidx >= qnum) qidx = 0;
            } else {
              const float x0 = RING[32 * ld], y0 = RING[32 * (2 + ld)];
              RING[32 * ld] = lr; RING[32 * (2 + ld)] = li;
              r0 = x0 * fr - y0 * fi;
              i0 = x0 * fi + y0 * fr;
#pragma unroll
              for (int m = 0; m < 3; m++) {
                const int o = 32 * (4 + (m == 0 ? 0 : m == 1 ? 3 : 7) + (m == 0 ? sd0 : m == 1 ? sd1 : sd2));
                const float x = RING[o], y = RING[o + 32 * 12];
                float re = x * sr[m] - y * si[m];
                float im = x * si[m] + y * sr[m];
                re += (-c[m]) * r0;
                im += (-c[m]) * i0;
                RING[o] = r0 + c[m] * re;
                RING[o + 32 * 12] = i0 + c[m] * im;
                r0 = re; i0 = im;
              }
            }
            const float tr = POW[k * kPowStride + bin];
            float rr = tr * r0, ri = tr * i0;
            if (++ld >= 2) ld = 0;
            if (++sd0 >= d0) sd0 = 0;
            if (++sd1 >= d1) sd1 = 0;
            if (++sd2 >= d2) sd2 = 0;
            H.step();
            H.apply(lr, li, rr, ri);
            if (pass == 0) {
              HL_re[k * kHyStride + sb] = lr; HL_im[k * kHyStride + sb] = li;
              HR_re[k * kHyStride + sb] = rr; HR_im[k * kHyStride + sb] = ri;
            } else {
              out_l[k * 128 + sb] = lr; out_l[k * 128 + 64 + sb] = li;
              out_r[k * 128 + sb] = rr; out_r[k * 128 + 64 + sb] = ri;
            }
          }
        }
        if (plain) {
          ist[4 + sb] = qidx;
          for (int r = 0; r < qnum; r++) { g_d[r * ds] = RING[32 * r]; g_d[d_im + r * ds] = RING[32 * (14 + r)]; }
        } else {
          for (int r = 0; r < 2; r++) { g_d[r * ds] = RING[32 * r]; g_d[d_im + r * ds] = RING[32 * (2 + r)]; }
#pragma unroll
          for (int m = 0; m < 3; m++)
#pragma unroll
            for (int r = 0; r < 3 + m; r++) {
              g_s[(m * 5 + r) * ds] = RING[32 * (4 + (m == 0 ? 0 : m == 1 ? 3 : 7) + r)];
              g_s[s_im + (m * 5 + r) * ds] = RING[32 * (16 + (m == 0 ? 0 : m == 1 ? 3 : 7) + r)];
            }
        }
        l_delay_end = ld; ser_end[0] = sd0; ser_end[1] = sd1; ser_end[2] = sd2;
      }
      if (pass == 0) {  // sub-subbands 4 and 5 were folded into 3 and 2; the right matrix keeps zeros there
        HR_re[lane * kHyStride + 4] = 0.f; HR_im[lane * kHyStride + 4] = 0.f;
        HR_re[lane * kHyStride + 5] = 0.f; HR_im[lane * kHyStride + 5] = 0.f;
      }
    }
    __syncwarp();
    if (lane == 0) { ist[0] = l_delay_end; ist[1] = ser_end[0]; ist[2] = ser_end[1]; ist[3] = ser_end[2]; }
    // ---- hybrid synthesis, lane = slot (ps_dec_flt.c:203-228): bands 0..2 = sums of 8, 2, 2 sub-subbands
    {
      const int cnt[3] = {8, 2, 2}, off[3] = {0, 8, 10};
      for (int b = 0; b < 3; b++) {
        float lr = 0.f, li = 0.f, rr = 0.f, ri = 0.f;
        for (int q = 0; q < cnt[b]; q++) {
          lr += HL_re[lane * kHyStride + off[b] + q]; li += HL_im[lane * kHyStride + off[b] + q];
          rr += HR_re[lane * kHyStride + off[b] + q]; ri += HR_im[lane * kHyStride + off[b] + q];
        }
        out_l[lane * 128 + b] = lr; out_l[lane * 128 + 64 + b] = li;
        out_r[lane * 128 + b] = rr; out_r[lane * 128 + 64 + b] = ri;
      }
    }
    __syncwarp();
  }
}

cudaError_t launch_esbr_ps(const EsbrPsArgs &args, int num_sms, cudaStream_t stream) {
  static PerDeviceOnce configured;
  const size_t smem = (size_t)(kFpsRomWords + kPsWarps * kPsWarpWords) * sizeof(float);
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(esbr_ps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.done();
  }
  long long need = (args.n_units + kPsWarps - 1) / kPsWarps;
  long long grid = (long long)num_sms * 2;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  esbr_ps_kernel<<<(unsigned)grid, kPsWarps * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace xb
