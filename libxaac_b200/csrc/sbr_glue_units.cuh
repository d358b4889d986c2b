// sbr_glue_units.cuh — the per-unit bodies of the block-floating-point bookkeeping of the fixed-point HQ SBR stage
// (ixheaacd_sbr_dec, decoder/ixheaacd_sbr_dec.c:662-1310, fixed branch, low_pow_flag = 0), one warp per unit.
//
// They are shared by the stand-alone glue kernels (sbr_glue_kernels.cu) and by the heavy kernels that absorb them:
//   sbr_pre_unit + sbr_scale_unit  -> sbr_front_hq_kernel (qmf_anal_kernel.cu): overlap rows, analysis bank, headroom
//                                     scans and rescale of one unit by one warp, the unit's rows still in L2
//   sbr_post_unit                  -> calc_sbrenvelope_hq_kernel (envcalc_kernel.cu), after the unit's envelope adjustment
// All matrix accesses are row segments (lanes = consecutive bands), i.e. coalesced 128-byte requests.
#pragma once
#include "fixmath.cuh"
#include "kernels.h"
#include "sbr_common.cuh"

namespace xb {

// sbr_dec.c:749-774: overlap slots -> matrix rows 0..5, ixheaacd_rescale_x_overlap (decoder/ixheaacd_sbrdec_lpfuncs.c:453-527).
// Returns the analysis bank's usb (qmf_bank->usb) of this frame.
XB_DEV int sbr_pre_unit(const SbrStageArgs &p, long long u, int lane) {
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
  int usb = misc[kMiscCodecUsb];
  if (p.side[u * kSideWords + kSideApply]) {
    const int old_lsb = misc[kMiscMaxQmfPrev], new_lsb = env[kEnvMaxQmfSubband];
    const int start_slot = env[kEnvTimeStep] * (misc[kMiscEndPosPrev] - env[kEnvNumTimeSlots]);
    const int syn_usb = misc[kMiscSynUsb];
    int ov_lb = sf[kSfOvLb], ov_hb = sf[kSfOvHb];
    __syncwarp();
    usb = (int16_t)new_lsb;
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
  return usb;
}

// sbr_dec.c:1050-1127: headroom scans (ixheaacd_expsubbandsamples), the three ixheaacd_adjust_scale calls, scale-factor
// update, ixheaacd_clr_subsamples; builds the HF generator's argument record.  `cur_mask` < 0: scan rows 6..37 here;
// otherwise it is the OR of ixheaac_abs32_nrm over rows 6..37 x bands < usb collected by the analysis bank of this warp.
XB_DEV void sbr_scale_unit(const SbrStageArgs &p, long long u, int lane, int usb, i32 cur_mask) {
  i32 *m = p.matrix + u * kSbrMatWords;
  i32 *lpc = p.lpc + u * 256;
  int16_t *sf = p.sf + u * 8;
  const int16_t *misc = p.misc + u * 16;
  const int16_t *side = p.side + u * kSideWords;
  int reserve = cur_mask < 0 ? warp_headroom(m, 0, usb, 6, 38, lane)
                             : pnorm32((i32)__reduce_or_sync(0xffffffffu, (unsigned)(cur_mask | 1)));
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
    const bool act = lane < usb && sh != 0;
#pragma unroll 1
    for (int l0 = 6; l0 < 38; l0 += 8) {  // eight rows per pass, their loads in flight together
      i32 a[8], b[8];
      if (act) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
          a[q] = m[128 * (l0 + q) + lane];
          b[q] = m[128 * (l0 + q) + 64 + lane];
        }
      }
#pragma unroll
      for (int q = 0; q < 8; q++) {
        i32 *row = m + 128 * (l0 + q);
        if (act) {
          row[lane] = sh > 0 ? lsl(a[q], sh) : (a[q] >> -sh);
          row[64 + lane] = sh > 0 ? lsl(b[q], sh) : (b[q] >> -sh);
        }
        row[32 + lane] = 0;
        row[96 + lane] = 0;
      }
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

// sbr_dec.c:1205-1245, :1284-1308: previous-frame data, LPC state rows, overlap save, synthesis parameters.
// `failed`: the envelope adjuster returned an error for this unit (the reference returns before any of this, :1203).
XB_DEV void sbr_post_unit(const SbrStageArgs &p, long long u, int lane, bool failed) {
  const i32 *m = p.matrix + u * kSbrMatWords;
  int16_t *sf = p.sf + u * 8, *misc = p.misc + u * 16;
  const int16_t *side = p.side + u * kSideWords;
  const int16_t *env = side + kSideEnv, *hfs = side + kSideHf;
  int16_t *synp = p.synp + u * 8;
  if (failed) {
    if (lane < 8) synp[lane] = 0;
    return;
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

}  // namespace xb
