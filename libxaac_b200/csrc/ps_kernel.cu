// ps_kernel.cu — fixed-point parametric stereo (hybrid analysis, decorrelation, rotation) for sm_100a (B200).
//
// One warp owns one unit (one HE-AACv2 frame).  The 32 QMF slots of the frame are processed in order (the delay lines
// and the interpolated mixing matrices carry from slot to slot); inside a slot the lanes own hybrid sub-subbands, QMF
// bands, parameter bins or stereo groups.  The whole PS state of the unit (5.2 KB) lives in shared memory for the frame.
// Replaces, bit-exactly, the PS work of the left ixheaacd_cplx_synt_qmffilt call (decoder/ixheaacd_qmf_dec.c:811-1129,
// active = 1) and its caller (decoder/ixheaacd_sbr_dec.c:1247-1262):
//   ixheaacd_init_ps_scale / ixheaacd_get_ps_scale / ixheaacd_scale_ps_states   ps_dec.c:125-210, thumb_ps_dec.c:101-181
//   the pre-shifts of the left call (ixheaacd_adjust_scale, qmf_dec.c:942-957)
//   ixheaacd_init_rot_env                                                        ps_dec.c:714-854
//   ixheaacd_apply_ps                                                            thumb_ps_dec.c:69-99
//     ixheaacd_hybrid_analysis, ixheaacd_filt_2_ch, ixheaacd_filt_8_ch           hybrid.c:51-285
//     ixheaacd_inv_dit_fft_8pt_dec (selector: ixheaacd_inv_dit_fft_8pt)          dsp_fft32x32s.c:34-117
//     ixheaacd_decorrelation_dec, ixheaacd_decorr_filter1_dec, _filter2_dec, ixheaacd_divide16_pos_dec (selector leaves)
//                                                                                ps_dec.c:212-675
//     ixheaacd_apply_rot_dec (selector: ixheaacd_apply_rot)                      ps_dec.c:856-991
//   ixheaacd_shiftrountine                                                       generic/ixheaacd_qmf_dec_generic.c:1610-1636
// Output: the left matrix in place (rows 0..31, already in the synthesis scale) and the right matrix, plus the
// parameter rows for the two ixheaacd_cplx_synt_qmffilt kernels that follow.
// Algorithmic HBM bytes per unit: 16 KB left rows in + 16 KB left out + 16 KB right out + 2 x 5.2 KB state + 1 KB
// parameters ~= 59.5 KB.
#include <cstdint>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"

namespace xb {

constexpr int kPsWarps = 4;

struct PsWarpS {
  int16_t st[kPsDspWords];  // PS state, blob layout (kernels.h kPsSt*)
  i32 rowL[128], rowR[128]; // the slot being processed: re[64] | im[64]
  i32 hyb[64];              // left_re[16] | left_im[16] | right_re[16] | right_im[16]
  int16_t tr[24];           // transient ratio per bin (+ tr[20] = 0)
};

XB_DEV i32 m16(i32 a, i32 b) { return a * b; }                                  // mult16x16in32
XB_DEV i32 m16_shl(i32 a, i32 b) { return sext16((a * b) >> 15); }              // mult16_shl
XB_DEV i32 rot_re(i32 r, i32 i, const int16_t *f) { return sext16(sub_sat(m16(r, f[0]), m16(i, f[1])) >> 15); }
XB_DEV i32 rot_im(i32 r, i32 i, const int16_t *f) { return sext16(add_sat(m16(r, f[1]), m16(i, f[0])) >> 15); }
XB_DEV i32 pw(i32 v) { return mul32x16(v, v >> 16); }
XB_DEV i32 mshl(i32 a, i32 c) { return lsl(mul32x16(a, c), 1); }                // shl32(mult32x16in32(a, c), 1)

// dsp_fft32x32s.c:34-117
XB_DEV void fft8(const i32 *y, i32 *real, i32 *imag) {
  i32 a0, a1, a2, a3, a00, a10, a20, a30, vr, vi, x[16];
  a00 = add_sat(y[0], y[8]); a0 = sub_sat(y[0], y[8]);
  a20 = add_sat(y[1], y[9]); a3 = sub_sat(y[1], y[9]);
  a10 = add_sat(y[4], y[12]); a2 = sub_sat(y[4], y[12]);
  a30 = add_sat(y[5], y[13]); a1 = sub_sat(y[5], y[13]);
  x[0] = add_sat(a00, a10); x[4] = sub_sat(a00, a10);
  x[1] = add_sat(a20, a30); x[5] = sub_sat(a20, a30);
  x[2] = sub_sat(a0, a1); x[6] = add_sat(a0, a1);
  x[3] = add_sat(a3, a2); x[7] = sub_sat(a3, a2);
  a00 = add_sat(y[2], y[10]); a0 = sub_sat(y[2], y[10]);
  a20 = add_sat(y[3], y[11]); a3 = sub_sat(y[3], y[11]);
  a10 = add_sat(y[6], y[14]); a2 = sub_sat(y[6], y[14]);
  a30 = add_sat(y[7], y[15]); a1 = sub_sat(y[7], y[15]);
  x[8] = add_sat(a00, a10); x[12] = sub_sat(a00, a10);
  x[9] = add_sat(a20, a30); x[13] = sub_sat(a20, a30);
  x[10] = sub_sat(a0, a1); x[14] = add_sat(a0, a1);
  x[11] = add_sat(a3, a2); x[15] = sub_sat(a3, a2);
  real[0] = add_sat(x[0], x[8]);
  imag[0] = add_sat(x[1], x[9]);
  a00 = sub_sat(x[0], x[8]);
  a10 = sub_sat(x[1], x[9]);
  a0 = sub_sat(x[4], x[13]);
  a1 = add_sat(x[5], x[12]);
  real[4] = add_sat(x[4], x[13]);
  imag[4] = sub_sat(x[5], x[12]);
  vr = mshl(sub_sat(x[10], x[11]), 0x5A82);
  vi = mshl(add_sat(x[10], x[11]), 0x5A82);
  real[1] = add_sat(x[2], vr);
  imag[1] = add_sat(x[3], vi);
  a2 = sub_sat(x[2], vr);
  a3 = sub_sat(x[3], vi);
  real[2] = add_sat(a0, a2);
  imag[2] = add_sat(a1, a3);
  vr = mshl(add_sat(x[14], x[15]), 0x5A82);
  vi = mshl(sub_sat(x[14], x[15]), 0x5A82);
  a20 = sub_sat(x[6], vr);
  a30 = add_sat(x[7], vi);
  real[3] = add_sat(a00, a20);
  imag[3] = add_sat(a10, a30);
  real[5] = add_sat(x[6], vr);
  imag[5] = sub_sat(x[7], vi);
}

// hybrid.c:96-212
XB_DEV void filt_8_ch(const i32 *re, const i32 *im, i32 *hr, i32 *hi, const int16_t *p) {
  const i32 tcos = 0x7642, tsin = 0x30fc, tcom = 0x5a82;
  i32 real, imag, cum[16];
#define MM(a, c) mul32x16((a), (c))
  real = lsl(add_sat(MM(re[0], p[0]), MM(re[8], p[8])), 1);
  imag = lsl(add_sat(MM(im[0], p[0]), MM(im[8], p[8])), 1);
  cum[12] = mshl(add_sat(imag, real), tcom);
  cum[13] = mshl(sub_sat(imag, real), tcom);
  real = lsl(add_sat(MM(re[1], p[1]), MM(re[9], p[9])), 1);
  imag = lsl(add_sat(MM(im[1], p[1]), MM(im[9], p[9])), 1);
  cum[10] = lsl(add_sat(MM(imag, tcos), MM(real, tsin)), 1);
  cum[11] = lsl(sub_sat(MM(imag, tsin), MM(real, tcos)), 1);
  cum[9] = mshl(sub_sat(re[2], re[10]), p[10]);
  cum[8] = mshl(sub_sat(im[2], im[10]), p[2]);
  real = lsl(add_sat(MM(re[3], p[3]), MM(re[11], p[11])), 1);
  imag = lsl(add_sat(MM(im[3], p[3]), MM(im[11], p[11])), 1);
  cum[6] = lsl(sub_sat(MM(imag, tcos), MM(real, tsin)), 1);
  cum[7] = lsl(neg_sat(add_sat(MM(imag, tsin), MM(real, tcos))), 1);
  real = lsl(add_sat(MM(re[4], p[4]), MM(re[12], p[12])), 1);
  imag = lsl(add_sat(MM(im[4], p[4]), MM(im[12], p[12])), 1);
  cum[4] = mshl(sub_sat(imag, real), tcom);
  cum[5] = mshl(neg_sat(add_sat(imag, real)), tcom);
  real = mshl(re[5], p[5]);
  imag = mshl(im[5], p[5]);
  cum[2] = lsl(sub_sat(MM(real, tcos), MM(imag, tsin)), 1);
  cum[3] = lsl(add_sat(MM(real, tsin), MM(imag, tcos)), 1);
  cum[0] = mshl(re[6], p[6]);
  cum[1] = mshl(im[6], p[6]);
  real = mshl(re[7], p[7]);
  imag = mshl(im[7], p[7]);
  cum[14] = lsl(add_sat(MM(imag, tsin), MM(real, tcos)), 1);
  cum[15] = lsl(sub_sat(MM(imag, tcos), MM(real, tsin)), 1);
#undef MM
  fft8(cum, hr, hi);
}

// hybrid.c:51-94, one component
XB_DEV void filt_2_ch(const i32 *q, i32 *h, const int16_t *p2_6) {
  const i32 cum0 = q[6] >> 1;
  i32 cum1 = 0;
#pragma unroll
  for (int j = 0; j < 6; j++) cum1 = add_sat(cum1, mul32x16(q[1 + 2 * j], p2_6[j]));
  cum1 = lsl(cum1, 1);
  h[0] = add_sat(cum0, cum1);
  h[1] = sub_sat(cum0, cum1);
}

// ps_dec.c:212-234
XB_DEV i32 divide16_pos(i32 op1, i32 op2) {
  const int nrm = norm32(op2);
  u32 u = (u32)op1 << nrm, v = (u32)op2 << nrm;
  u &= 0xffff0000u;
  v &= 0xffff0000u;
  if (u != 0) {
#pragma unroll 1
    for (int k = 16; k > 0; k--) {
      if (u >= v) u = ((u - v) << 1) + 1;
      else u <<= 1;
    }
  }
  return (i32)u;
}

// ps_dec.c:677-712
XB_DEV i32 cos512(i32 phi, const int16_t *tab) {
  const i32 a = phi == (i32)0x80000000 ? 0x7fffffff : (phi < 0 ? -phi : phi);
  const int index = round16(a) & 0x3ff;
  return index < 512 ? tab[512 - index] : sext16(-tab[index - 512]);
}
XB_DEV i32 sin512(i32 phi, const int16_t *tab) {
  int index = round16(phi);
  if (index < 0) {
    index = (-index) & 0x3ff;
    return index < 512 ? sext16(-tab[index]) : sext16(-tab[1024 - index]);
  }
  index &= 0x3ff;
  return index < 512 ? tab[index] : tab[1024 - index];
}

XB_DEV i32 blockshift(i32 v, int sh) {  // ixheaacd_adjust_scale_dec semantics (env_calc.c:1099-1157)
  if (sh == 0) return v;
  sh = max(-31, min(31, sh));
  return sh > 0 ? lsl(v, sh) : (v >> -sh);
}

__global__ void __launch_bounds__(kPsWarps * 32) ps_frame_kernel(PsArgs p) {
  __shared__ PsWarpS ws[kPsWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;
  PsWarpS &w = ws[warp];
  const int16_t *rom = reinterpret_cast<const int16_t *>(p.ps_rom);
  const int16_t *trig = reinterpret_cast<const int16_t *>(p.misc_rom);
  const int16_t *inv_int = reinterpret_cast<const int16_t *>(p.env_rom + kERomInvInt);
  const int16_t *borders = rom + kPsRomBordersGroup;
  const long long warps_total = (long long)gridDim.x * kPsWarps;

  for (long long u = (long long)blockIdx.x * kPsWarps + warp; u < p.n_units; u += warps_total) {
    const int16_t *side = p.side + u * kSideWords;
    const int mode = side[kSidePs];
    if (mode == 0 || !side[kSideApply] || side[kSideEnv + kEnvChannelMode] != 3 || (p.err && p.err[u] != 0)) {
      if (lane == 0) p.ps_done[u] = 0;  // gates the right-channel synthesis of this unit
      continue;
    }
    const bool as_built = mode != 2;
    const int16_t *prm = side + kSidePsPrm;
    int16_t *st = w.st;
    __syncwarp();
    {
      const i32 *src = reinterpret_cast<const i32 *>(p.ps_state + u * kPsDspWords);
      for (int i = lane; i < kPsDspWords / 2; i += 32) reinterpret_cast<i32 *>(st)[i] = src[i];
      w.hyb[lane] = 0;
      w.hyb[32 + lane] = 0;
    }
    __syncwarp();
    int16_t *idx = st + kPsStIdx;
    i32 *peak = reinterpret_cast<i32 *>(st + kPsStPeak), *hybq = reinterpret_cast<i32 *>(st + kPsStHyb);
    int16_t *synp = p.synp + u * 8, *synp_r = p.synp_r + u * 8;
    int16_t *sf = p.sf + u * 8, *sf_r = p.sf_r + u * 8;
    const int ov_lb = synp[0], lb = synp[1], hb = synp[2], st_syn = synp[3], lsb = synp[4], usb = synp[5];

    // ---- ixheaacd_init_ps_scale ----
    int ps_scale;
    {
      i32 mx = 0;
      auto scan16 = [&](const int16_t *q, int n) {
        for (int i = lane; i < n; i += 32) mx |= abs_nrm((i32)q[i]);
      };
      scan16(st + kPsStAp + 6, 40); scan16(st + kPsStAp + 64 + 6, 40);
      scan16(st + kPsStLd, 336); scan16(st + kPsStSd, 58); scan16(st + kPsStSub, 64); scan16(st + kPsStSubSer, 480);
      for (int i = 0; i < 3; i++)
        for (int m = 0; m < rom[kPsRomRevDelay + i]; m++) scan16(st + kPsStSer + 192 * m + 64 * i + 6, 40);
      mx = (i32)((u32)mx << 16);
      for (int i = lane; i < 72; i += 32) mx |= abs_nrm(hybq[i]);
      mx = (i32)__reduce_or_sync(full, (unsigned)mx);
      const int reserve = (mx == 0) ? 31 : pnorm32(mx);
      const int dscale = sext16(idx[kPsIdxScale] + reserve);
      int t = min(min(lb, ov_lb), min(hb, dscale));
      ps_scale = sext16(t - 1);
      const int scale = sext16((ps_scale - dscale) + reserve);
      __syncwarp();
      auto sh16 = [&](int16_t *q, int n) {
        if (scale > 0) {
          const int s1 = min(scale, 15);
          for (int i = lane; i < n; i += 32) q[i] = (int16_t)sat16((i32)q[i] << s1);
        } else {
          const int s1 = min(-scale, 31);
          for (int i = lane; i < n; i += 32) q[i] = (int16_t)((i32)q[i] >> s1);
        }
      };
      if (scale != 0) {
        sh16(st + kPsStAp + 6, 40); sh16(st + kPsStAp + 64 + 6, 40);
        sh16(st + kPsStLd, 336); sh16(st + kPsStSd, 58); sh16(st + kPsStSub, 64); sh16(st + kPsStSubSer, 480);
        for (int i = 0; i < 3; i++)
          for (int m = 0; m < rom[kPsRomRevDelay + i]; m++) sh16(st + kPsStSer + 192 * m + 64 * i + 6, 40);
        const int s2 = sext16(scale > 0 ? scale + scale : -(scale + scale));
        for (int i = lane; i < 72; i += 32) hybq[i] = scale > 0 ? shl32_sat(hybq[i], min(scale, 31)) : shr32(hybq[i], -scale);
        for (int i = lane; i < 60; i += 32) peak[i] = scale > 0 ? shl32_sat(peak[i], min(s2, 31)) : shr32(peak[i], s2);
      }
      __syncwarp();
      if (lane == 0) idx[kPsIdxScale] = (int16_t)ps_scale;
    }
    const int ov_lb_shift = ps_scale - ov_lb, lb_shift = ps_scale - lb, hb_shift = ps_scale - hb;
    const int common_shift = (st_syn - ps_scale) - 8;
    const int shiftdelay_late = sext16(lb - ps_scale);
    // pre-shift of the left call (qmf_dec.c:942-957) for (band k, slot l < 32)
    auto pre = [&](int k, int l) { return k < lsb ? (l < 6 ? ov_lb_shift : lb_shift) : (k < usb ? hb_shift : 0); };

    i32 *mat = p.matrix + u * kSbrMatWords;
    i32 *right = p.right + u * 4096;
    int env = 0;
    int ps_usb = idx[kPsIdxUsb];
    int d_idx = idx[kPsIdxDelay], d_long = idx[kPsIdxDelayLong];
    int d_ser0 = idx[kPsIdxSer], d_ser1 = idx[kPsIdxSer + 1], d_ser2 = idx[kPsIdxSer + 2];
    int16_t *hv = st + kPsStHvec;
    int16_t *h11v = hv, *h21v = hv + 48, *H11 = hv + 96, *H21 = hv + 144, *d11 = hv + 192, *d21 = hv + 240;
    __syncwarp();

#pragma unroll 1
    for (int slot = 0; slot < 32; slot++) {
      // ---- ixheaacd_init_rot_env at PS envelope borders ----
      if (env < 7 && slot == prm[kPsPrmBorder + env]) {
        if (env == 0) {
          const int usb_prev = ps_usb;
          ps_usb = usb;
          if (usb > usb_prev && usb_prev) {
            const int o = min(usb, 20);
            if (o > usb_prev)
              for (int i = 0; i < 3; i++)
                for (int j = 0; j < rom[kPsRomRevDelay + i]; j++)
                  for (int q = lane; q < 2 * (o - usb_prev); q += 32) st[kPsStSer + 192 * j + 64 * i + 2 * usb_prev + q] = 0;
            const int o1 = min(usb, 32);
            if (o1 >= o && o1 <= 12)
              for (int i = 0; i < 14; i++)
                for (int q = lane; q < 2 * (o1 - o); q += 32) st[kPsStLd + 24 * i + 2 * o + q] = 0;
            if (usb >= o1 && usb <= 16)
              for (int q = lane; q < 2 * (usb - o1); q += 32) st[kPsStSd + 2 * o1 + q] = 0;
          }
        }
        if (lane < 22) {
          const int g = lane;
          const bool fine = prm[kPsPrmIidQuant] != 0;
          const int steps = fine ? 15 : 7;
          const int16_t *sfac = rom + (fine ? kPsRomScaleFine : kPsRomScale);
          const i32 dl = sat16(prm[kPsPrmBorder + env + 1] - prm[kPsPrmBorder + env]);
          const i32 inv_len = inv_int[sext16(dl < 0 ? -dl : dl)];
          const int bin = rom[kPsRomGroupToBin + g];
          const int ii = prm[kPsPrmIid + 34 * env + bin], ic = prm[kPsPrmIcc + 34 * env + bin];
          const i32 c1 = sfac[steps + ii], c2 = sfac[steps - ii];
          const i32 al = rom[kPsRomAlpha + ic];
          const i32 beta = lsl(mul32x16(shl32(al * sext16(c1 - c2), 1), 0x5a82), 1);
          const i32 alpha = (al << 16) >> 1;
          const i32 bpa = round16(add_sat(beta, alpha)), bma = round16(sub_sat(beta, alpha));
          const i32 rescale = (i32)(0x0517cc1bu << 1);
          const i32 ipa = mul32x16(rescale, bpa), ima = mul32x16(rescale, bma);
          const i32 h11 = m16_shl(cos512(ipa, trig), c2), h12 = m16_shl(cos512(ima, trig), c1);
          const i32 h21 = m16_shl(sin512(ipa, trig), c2), h22 = m16_shl(sin512(ima, trig), c1);
          d11[2 * g] = (int16_t)m16_shl(inv_len, sext16(h11 - h11v[2 * g]));
          d11[2 * g + 1] = (int16_t)m16_shl(inv_len, sext16(h12 - h11v[2 * g + 1]));
          d21[2 * g] = (int16_t)m16_shl(inv_len, sext16(h21 - h21v[2 * g]));
          d21[2 * g + 1] = (int16_t)m16_shl(inv_len, sext16(h22 - h21v[2 * g + 1]));
          H11[2 * g] = h11v[2 * g]; H11[2 * g + 1] = h11v[2 * g + 1];
          H21[2 * g] = h21v[2 * g]; H21[2 * g + 1] = h21v[2 * g + 1];
          h11v[2 * g] = (int16_t)h11; h11v[2 * g + 1] = (int16_t)h12;
          h21v[2 * g] = (int16_t)h21; h21v[2 * g + 1] = (int16_t)h22;
        }
        env++;
        __syncwarp();
      }

      // ---- load the left row (pre-shifted) ----
      {
        const i32 *row = mat + 128 * slot;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int k = lane + 32 * h, sh = pre(k, slot);
          w.rowL[k] = blockshift(row[k], sh);
          w.rowL[64 + k] = blockshift(row[64 + k], sh);
        }
      }
      // ---- hybrid analysis of slot + 6 (hybrid.c:214-285): lane b < 3 owns QMF band b ----
      if (lane < 3) {
        const int band = lane, l6 = slot + 6;
        i32 wre[13], wim[13];
        i32 *bre = hybq + 24 * band, *bim = bre + 12;
#pragma unroll
        for (int j = 0; j < 12; j++) { wre[j] = bre[j]; wim[j] = bim[j]; }
        i32 tr_ = mat[128 * l6 + band], ti_ = mat[128 * l6 + 64 + band];
        if (l6 < 32) { const int sh = pre(band, l6); tr_ = blockshift(tr_, sh); ti_ = blockshift(ti_, sh); }
        const int sd = slot < 26 ? 0 : shiftdelay_late;
        if (sd < 0) { tr_ = shl32(tr_, -sd); ti_ = shl32(ti_, -sd); }
        else { tr_ = shr32(tr_, sd); ti_ = shr32(ti_, sd); }
        wre[12] = tr_;
        wim[12] = ti_;
#pragma unroll
        for (int j = 0; j < 12; j++) { bre[j] = wre[j + 1]; bim[j] = wim[j + 1]; }
        if (band == 0) {
          filt_8_ch(wre, wim, w.hyb, w.hyb + 16, rom + kPsRomP8_13);
        } else {
          const int off = 6 + 2 * (band - 1);
          filt_2_ch(wre, w.hyb + off, rom + kPsRomP2_6);
          filt_2_ch(wim, w.hyb + 16 + off, rom + kPsRomP2_6);
        }
      }
      __syncwarp();
      const i32 *lre = w.hyb, *lim = w.hyb + 16;
      i32 *rre = w.hyb + 32, *rim = w.hyb + 48;

      // ---- power per bin and transient ratio (ps_dec.c:482-590): lane = bin ----
      if (lane < 20) {
        const int bin = lane;
        i32 pwr;
        if (bin == 0) pwr = add_sat(add_sat(add_sat(pw(lre[0]), pw(lim[0])), pw(lre[5])), pw(lim[5]));
        else if (bin == 1) pwr = add_sat(add_sat(add_sat(pw(lre[4]), pw(lim[4])), pw(lre[1])), pw(lim[1]));
        else if (bin < 8) { const int sb = borders[bin + 2]; pwr = add_sat(pw(lre[sb]), pw(lim[sb])); }
        else if (bin < 14) { const int sb = bin - 5; pwr = add_sat(pw(w.rowL[sb]), pw(w.rowL[64 + sb])); }
        else {
          const int gr = bin + 2;
          const int mxs = min(ps_usb, (int)borders[gr + 1]), gs = rom[kPsRomGroupShift + gr - 16];
          pwr = 0;
          for (int sb = borders[gr]; sb < mxs; sb++)
            pwr = add_sat(pwr, add_sat(pw(w.rowL[sb]), pw(w.rowL[64 + sb])) >> gs);
        }
        i32 pv = shl32(pwr, 1);
        if (pv < 0) pv = 0;
        i32 pk = lsl(mul32x16(peak[bin], 0x620a), 1);
        if (pv > pk) pk = pv;
        peak[bin] = pk;
        i32 pd = add_sat(lsl(mul32x16(peak[40 + bin], 0x6000), 1), sub_sat(pk, pv) >> 2);
        peak[40 + bin] = pd;
        const i32 nrg = add_sat(lsl(mul32x16(peak[20 + bin], 0x6000), 1), pv >> 2);
        peak[20 + bin] = nrg;
        pd = add_sat(pd, pd >> 1);
        w.tr[bin] = pd <= nrg ? (int16_t)0x7fff : (int16_t)divide16_pos(nrg, pd);
      } else if (lane == 20) {
        w.tr[20] = 0;
      }
      __syncwarp();

      // ---- all-pass decorrelators: lanes 0..9 hybrid sub-subbands, lanes 10..29 QMF bands 3..22 ----
      if (lane < 30) {
        const bool hy = lane < 10;
        const int sb = hy ? lane : lane - 7;
        int16_t *dl = hy ? st + kPsStSub + 32 * d_idx + 2 * sb : st + kPsStAp + 64 * d_idx + 2 * sb;
        const int16_t *fac = rom + (hy ? kPsRomFracSub : kPsRomFracQmf) + 2 * sb;
        const i32 r0 = dl[0], i0 = dl[1];
        i32 rin = rot_re(r0, i0, fac), iin = rot_im(r0, i0, fac);
        const i32 inr = hy ? lre[sb] : w.rowL[sb], ini = hy ? lim[sb] : w.rowL[64 + sb];
        dl[0] = (int16_t)round16(inr);
        dl[1] = (int16_t)round16(ini);
#pragma unroll
        for (int m = 0; m < 3; m++) {
          const int di = m == 0 ? d_ser0 : (m == 1 ? d_ser1 : d_ser2);
          int16_t *q = hy ? st + kPsStSubSer + 96 * di + 32 * m + 2 * sb : st + kPsStSer + 192 * di + 64 * m + 2 * sb;
          const int16_t *f = rom + (hy ? kPsRomFracSubSer + 32 * m : kPsRomFracQmfSer + 64 * m) + 2 * sb;
          const i32 decay = hy ? rom[kPsRomRevDecay + m] : rom[kPsRomDecaySf + 3 * sb + m];
          const i32 q0 = q[0], q1 = q[1];
          i32 rt = rot_re(q0, q1, f), it = rot_im(q0, q1, f);
          rt = sext16(rt - m16_shl(rin, decay));
          it = sext16(it - m16_shl(iin, decay));
          q[0] = (int16_t)(rin + m16_shl(rt, decay));
          q[1] = (int16_t)(iin + m16_shl(it, decay));
          rin = rt;
          iin = it;
        }
        const i32 t = w.tr[hy ? rom[kPsRomHybToBin + sb] : rom[kPsRomDelayToBin + sb]];
        const i32 outr = shl32(rin * t, 1), outi = shl32(iin * t, 1);
        if (hy) { rre[sb] = outr; rim[sb] = outi; }
        else { w.rowR[sb] = outr; w.rowR[64 + sb] = outi; }
      }
      __syncwarp();
      // ---- plain delays (ps_dec.c:596-645) and clearing above usb ----
      {
        const int b20 = borders[20], b21 = borders[21], b22 = borders[22];
        const int us = sext16(ps_usb);
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int k = lane + 32 * h;
          if (k >= b20 && k < min(us, b21)) {
            int16_t *d = st + kPsStLd + 24 * d_long + 2 * (k - b20);
            const i32 r = d[0], i = d[1], t = w.tr[18];
            d[0] = (int16_t)round16(w.rowL[k]);
            d[1] = (int16_t)round16(w.rowL[64 + k]);
            w.rowR[k] = shl32(r * t, 1);
            w.rowR[64 + k] = shl32(i * t, 1);
          } else if (k >= b21 && k < min(us, b22)) {
            int16_t *d = st + kPsStSd + 2 * (k - b21);
            const i32 r = d[0], i = d[1], t = w.tr[19];
            d[0] = (int16_t)round16(w.rowL[k]);
            d[1] = (int16_t)round16(w.rowL[64 + k]);
            w.rowR[k] = shl32(r * t, 1);
            w.rowR[64 + k] = shl32(i * t, 1);
          }
          if (k >= ps_usb) { w.rowR[k] = 0; w.rowR[64 + k] = 0; }
        }
      }
      d_long = sext16(d_long + 1);
      if (d_long >= 14) d_long = 0;
      d_idx = d_idx + 1 >= 2 ? 0 : d_idx + 1;
      d_ser0 = d_ser0 + 1 >= rom[kPsRomRevDelay] ? 0 : d_ser0 + 1;
      d_ser1 = d_ser1 + 1 >= rom[kPsRomRevDelay + 1] ? 0 : d_ser1 + 1;
      d_ser2 = d_ser2 + 1 >= rom[kPsRomRevDelay + 2] ? 0 : d_ser2 + 1;
      // ---- rotation (ps_dec.c:856-991) ----
      for (int j = lane; j < 44; j += 32) {
        H11[j] = (int16_t)(H11[j] + d11[j]);
        H21[j] = (int16_t)(H21[j] + d21[j]);
      }
      __syncwarp();
      if (lane < 10) {
        const int s = lane;
        const i32 a = add_sat(mul32x16(lre[s], H11[2 * s]), mul32x16(rre[s], H21[2 * s]));
        const i32 b = add_sat(mul32x16(lim[s], H11[2 * s]), mul32x16(rim[s], H21[2 * s]));
        const i32 c = add_sat(mul32x16(lre[s], H11[2 * s + 1]), mul32x16(rre[s], H21[2 * s + 1]));
        const i32 d = add_sat(mul32x16(lim[s], H11[2 * s + 1]), mul32x16(rim[s], H21[2 * s + 1]));
        w.hyb[s] = shl32(a, 2);
        w.hyb[16 + s] = shl32(b, 2);
        w.hyb[32 + s] = shl32(c, 2);
        w.hyb[48 + s] = shl32(d, 2);
      }
      // QMF bands 3..usb-1
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int k = lane + 32 * h;
        if (k >= 3 && k < ps_usb) {
          i32 h0 = 0, h1 = 0, h2 = 0, h3 = 0;
          if (!as_built) {  // see apply_rot in oracle/src/ps.c: the reference's x86-64 gcc build reads these as zero
            int g = 10;
            while (g < 21 && k >= borders[g + 1]) g++;
            if (k >= borders[g] && k < min(ps_usb, (int)borders[g + 1])) {
              h0 = H11[2 * g]; h1 = H11[2 * g + 1]; h2 = H21[2 * g]; h3 = H21[2 * g + 1];
            }
          }
          const i32 lr = w.rowL[k], li = w.rowL[64 + k], rr = w.rowR[k], ri = w.rowR[64 + k];
          w.rowL[k] = shl32(add_sat(mul32x16(lr, h0), mul32x16(rr, h2)), 2);
          w.rowL[64 + k] = shl32(add_sat(mul32x16(li, h0), mul32x16(ri, h2)), 2);
          w.rowR[k] = shl32(add_sat(mul32x16(lr, h1), mul32x16(rr, h3)), 2);
          w.rowR[64 + k] = shl32(add_sat(mul32x16(li, h1), mul32x16(ri, h3)), 2);
        }
      }
      __syncwarp();
      if (lane < 3) {  // fold the hybrid sub-subbands back into QMF bands 0..2 (resolutions 8->6, 2, 2)
        const int s = lane, o = s == 0 ? 0 : 6 + 2 * (s - 1), cnt = s == 0 ? 6 : 2;
        i32 a = w.hyb[o], b = w.hyb[16 + o], c = w.hyb[32 + o], d = w.hyb[48 + o];
        for (int q = 1; q < cnt; q++) {
          a = add_sat(a, w.hyb[o + q]);
          b = add_sat(b, w.hyb[16 + o + q]);
          c = add_sat(c, w.hyb[32 + o + q]);
          d = add_sat(d, w.hyb[48 + o + q]);
        }
        w.rowL[s] = a; w.rowL[64 + s] = b; w.rowR[s] = c; w.rowR[64 + s] = d;
      }
      __syncwarp();
      // ---- ixheaacd_shiftrountine on the left row, store both rows ----
      {
        i32 *lrow = mat + 128 * slot, *rrow = right + 128 * slot;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int k = lane + 32 * q;
          i32 v = w.rowL[k];
          if (common_shift < 0) v = v >> min(-common_shift, 31);
          else if (common_shift > 0) v = shl32_sat(v, min(common_shift, 31));
          lrow[k] = v;
          rrow[k] = w.rowR[k];
        }
      }
      __syncwarp();
    }

    if (lane == 0) {
      p.ps_done[u] = 1;
      idx[kPsIdxUsb] = (int16_t)ps_usb;
      idx[kPsIdxDelay] = (int16_t)d_idx;
      idx[kPsIdxDelayLong] = (int16_t)d_long;
      idx[kPsIdxSer] = (int16_t)d_ser0;
      idx[kPsIdxSer + 1] = (int16_t)d_ser1;
      idx[kPsIdxSer + 2] = (int16_t)d_ser2;
      sf[kSfPs] = (int16_t)ps_scale;
      sf_r[kSfOvLb] = sf_r[kSfLb] = sf_r[kSfHb] = (int16_t)ps_scale;  // sbr_dec.c:1259-1262
      // left: the matrix is already in the synthesis scale -> zero block shifts in the synthesis kernel
      synp[0] = synp[1] = synp[2] = (int16_t)(st_syn - 8);
      synp_r[0] = synp_r[1] = synp_r[2] = (int16_t)ps_scale;
      synp_r[3] = sf_r[kSfStSyn];
      synp_r[4] = idx[kPsIdxLsbR];
      synp_r[5] = idx[kPsIdxUsbR];
      synp_r[6] = 6;
      synp_r[7] = 0;
    }
    __syncwarp();
    {
      i32 *dst = reinterpret_cast<i32 *>(p.ps_state + u * kPsDspWords);
      for (int i = lane; i < kPsDspWords / 2; i += 32) dst[i] = reinterpret_cast<const i32 *>(st)[i];
    }
  }
}

cudaError_t launch_ps_frame(const PsArgs &args, int num_sms, cudaStream_t stream) {
  long long need = (args.n_units + kPsWarps - 1) / kPsWarps;
  long long grid = (long long)num_sms * 8;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  ps_frame_kernel<<<(unsigned)grid, kPsWarps * 32, 0, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace xb
