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
#include <type_traits>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"

namespace xb {

constexpr int kPsWarps = 4;

struct PsWarpS {
  int16_t st[kPsDspWords];  // PS state, blob layout (kernels.h kPsSt*)
  i32 hyb[64];              // left_re[16] | left_im[16] | right_re[16] | right_im[16]
  i32 xs[3][2][44];         // hybrid filter input history: 12 delayed + 32 new samples of QMF bands 0..2 (re, im)
  i32 hybL[32][21];         // hybrid analysis of all 32 slots: [slot][re 0..9 | im 10..19] (odd row stride: no conflicts)
  u32 pwt[32][21];          // power per parameter bin of all 32 slots: [slot][bin 0..19] (odd row stride)
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

// ps_dec.c:212-234, low 16 bits of the result.  The reference runs a 16-step restoring division on the normalised
// high halves U = (op1 << n) >> 16, V = (op2 << n) >> 16; for 0 <= op1 < op2 (the only way it is called, :574-579) the
// quotient bits land in the low half as floor(U * 2^15 / V) (first bit has weight 2^15), which is one 32-bit division.
XB_DEV i32 divide16_pos_lo(i32 op1, i32 op2) {
  const int nrm = norm32(op2);
  const u32 U = ((u32)op1 << nrm) >> 16, V = ((u32)op2 << nrm) >> 16;
  return (i32)(((U << 15) / V) & 0xffffu);
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

template <bool NOSAT>
__global__ void __launch_bounds__(kPsWarps * 32, 4) ps_frame_kernel(PsArgs p) {
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
    {  // pull this warp's next unit towards L2 while this one is processed
      const long long un = u + warps_total;
      if (un < p.n_units) {
        const char *q0 = reinterpret_cast<const char *>(p.ps_state + un * kPsDspWords);
        for (int o = lane * 128; o < kPsDspWords * 2; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q0 + o));
        const char *q1 = reinterpret_cast<const char *>(p.side + un * kSideWords);
        for (int o = lane * 128; o < kSideWords * 2; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q1 + o));
      }
    }
    __syncwarp();
    {
      const i32 *src = reinterpret_cast<const i32 *>(p.ps_state + u * kPsDspWords);
      // 649 8-byte words (the unit stride 5192 B is 8-byte aligned): all of a lane's 21 loads in flight before the stores
      static_assert(kPsDspWords % 4 == 0, "PS state blob must be a whole number of 8-byte words");
      const int2 *src2 = reinterpret_cast<const int2 *>(src);
      int2 v[21];
#pragma unroll
      for (int q = 0; q < 21; q++) {
        const int i = lane + 32 * q;
        v[q] = i < kPsDspWords / 4 ? src2[i] : make_int2(0, 0);
      }
#pragma unroll
      for (int q = 0; q < 21; q++) {
        const int i = lane + 32 * q;
        if (i < kPsDspWords / 4) {
          reinterpret_cast<i32 *>(st)[2 * i] = v[q].x;
          reinterpret_cast<i32 *>(st)[2 * i + 1] = v[q].y;
        }
      }
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

    // ---- hybrid analysis of all 32 slots (hybrid.c:214-285): a 13-tap FIR over time, so the slots are independent:
    //      lane = slot.  x[t], t = -12..31: 12 delayed samples from the state, then QMF rows 6..37 of bands 0..2 in
    //      the scale the reference sees them in (pre-shifted rows < 32, ixheaacd_apply_ps' shiftdelay for slots >= 26)
    for (int i = lane; i < 3 * 2 * 44; i += 32) {
      const int b = i / 88, c = (i / 44) & 1, j = i % 44;
      i32 v;
      if (j < 12) {
        v = hybq[24 * b + 12 * c + j];
      } else {
        const int t = j - 12, l6 = t + 6;
        v = mat[128 * l6 + 64 * c + b];
        if (l6 < 32) v = blockshift(v, pre(b, l6));
        const int sd = t < 26 ? 0 : shiftdelay_late;
        v = sd < 0 ? shl32(v, -sd) : shr32(v, sd);
      }
      w.xs[b][c][j] = v;
    }
    __syncwarp();
    {
      const int s_ = lane;
      i32 wre[13], wim[13], hr[6], hi[6];
#pragma unroll
      for (int j = 0; j < 13; j++) { wre[j] = w.xs[0][0][s_ + j]; wim[j] = w.xs[0][1][s_ + j]; }
      filt_8_ch(wre, wim, hr, hi, rom + kPsRomP8_13);
#pragma unroll
      for (int q = 0; q < 6; q++) { w.hybL[s_][q] = hr[q]; w.hybL[s_][10 + q] = hi[q]; }
#pragma unroll
      for (int b = 1; b < 3; b++) {
#pragma unroll
        for (int j = 0; j < 13; j++) { wre[j] = w.xs[b][0][s_ + j]; wim[j] = w.xs[b][1][s_ + j]; }
        filt_2_ch(wre, hr, rom + kPsRomP2_6);
        filt_2_ch(wim, hi, rom + kPsRomP2_6);
        w.hybL[s_][4 + 2 * b] = hr[0]; w.hybL[s_][5 + 2 * b] = hr[1];
        w.hybL[s_][14 + 2 * b] = hi[0]; w.hybL[s_][15 + 2 * b] = hi[1];
      }
    }
    __syncwarp();
    for (int i = lane; i < 72; i += 32) hybq[i] = w.xs[i / 24][(i / 12) & 1][32 + i % 12];
    __syncwarp();

    // ---- per-lane constants of the slot loop ----
    // all-pass decorrelators: lanes 0..9 hybrid sub-subbands, lanes 10..29 QMF bands 3..22
    const bool ap_act = lane < 30, hy = lane < 10;
    const int sb = hy ? lane : (ap_act ? lane - 7 : 3);
    int16_t *ap_dl = hy ? st + kPsStSub + 2 * sb : st + kPsStAp + 2 * sb;
    const int ap_dl_stride = hy ? 32 : 64;
    int16_t *ap_q = hy ? st + kPsStSubSer + 2 * sb : st + kPsStSer + 2 * sb;
    const int ap_q_stride = hy ? 96 : 192, ap_m_stride = hy ? 32 : 64;
    const int16_t *ff = rom + (hy ? kPsRomFracSub : kPsRomFracQmf) + 2 * sb;
    const i32 f0r = ff[0], f0i = ff[1];
    i32 fmr[3], fmi[3], dec[3];
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const int16_t *f = rom + (hy ? kPsRomFracSubSer + 32 * m : kPsRomFracQmfSer + 64 * m) + 2 * sb;
      fmr[m] = f[0];
      fmi[m] = f[1];
      dec[m] = hy ? rom[kPsRomRevDecay + m] : rom[kPsRomDecaySf + 3 * sb + m];
    }
    const int trbin = hy ? rom[kPsRomHybToBin + sb] : rom[kPsRomDelayToBin + sb];
    const int rd0 = rom[kPsRomRevDelay], rd1 = rom[kPsRomRevDelay + 1], rd2 = rom[kPsRomRevDelay + 2];
    // NOSAT: no fractional-delay factor of the ROM is -32768, the 16x16 rotations cannot saturate (one kernel per case: the
    // saturating forms cost 5 instructions per add and were 7 % of the kernel when selected at run time)
    auto rotr = [&](i32 r, i32 i, i32 fr, i32 fi) { return sext16((NOSAT ? r * fr - i * fi : sub_sat(r * fr, i * fi)) >> 15); };
    auto roti = [&](i32 r, i32 i, i32 fr, i32 fi) { return sext16((NOSAT ? r * fi + i * fr : add_sat(r * fi, i * fr)) >> 15); };
    // mixing matrices: lane g < 22 keeps h11/h12/h21/h22 of stereo group g (previous envelope target, interpolated
    // value, per-slot increment) in registers for the whole frame
    const int gi = lane < 22 ? lane : 0;
    i32 hv11 = h11v[2 * gi], hv12 = h11v[2 * gi + 1], hv21 = h21v[2 * gi], hv22 = h21v[2 * gi + 1];
    i32 H11r = H11[2 * gi], H12r = H11[2 * gi + 1], H21r = H21[2 * gi], H22r = H21[2 * gi + 1];
    i32 D11r = d11[2 * gi], D12r = d11[2 * gi + 1], D21r = d21[2 * gi], D22r = d21[2 * gi + 1];
    const int b20 = borders[20], b21 = borders[21], b22 = borders[22];
    // pre-shifts of this lane's two bands as (mul, shr) pairs: value * mul >> shr, branch-free in the slot loop
    i32 mA_ov, mA_lb, mB_ov, mB_lb;
    int rA_ov, rA_lb, rB_ov, rB_lb;
    {
      auto enc = [](int sh, i32 &mul, int &shr) {
        sh = max(-31, min(31, sh));
        mul = sh > 0 ? (i32)(1u << sh) : 1;
        shr = sh < 0 ? -sh : 0;
      };
      enc(pre(lane, 0), mA_ov, rA_ov);
      enc(pre(lane, 6), mA_lb, rA_lb);
      enc(pre(lane + 32, 0), mB_ov, rB_ov);
      enc(pre(lane + 32, 6), mB_lb, rB_lb);
    }

    // ---- power per parameter bin of all 32 slots (ps_dec.c:482-556): it does not depend on the slot-serial peak state, so
    //      lane = slot sums its own row serially instead of six warp reductions per slot.  All terms are >= 0, so the
    //      reference's running add32_sat equals min(MAX_32, exact sum); the exact group sums stay below 2^32 after the
    //      group shifts.  Bands below the first PS border slot still use the previous frame's usb.
    {
      const int s_ = lane;
      const i32 *row = mat + 128 * s_;
      const int border0 = prm[kPsPrmBorder];
      const int usb_eff = s_ >= border0 ? usb : ps_usb;
      // block shifts of this lane's slot as (mul, shr) pairs per band region: value * mul >> shr, branch-free
      auto enc = [](int sh, i32 &mul, int &shr) {
        sh = max(-31, min(31, sh));
        mul = sh > 0 ? (i32)(1u << sh) : 1;
        shr = sh < 0 ? -sh : 0;
      };
      i32 mul_lo, mul_hb;
      int shr_lo, shr_hb;
      enc(s_ < 6 ? ov_lb_shift : lb_shift, mul_lo, shr_lo);
      enc(hb_shift, mul_hb, shr_hb);
      // The lane walks its own row in 16-byte requests (four bands of one component each: 32 requests instead of 122 scalar
      // ones that each touched one word of 32 different sectors — 18 % of the kernel's stall samples sat on them).  The band ->
      // bin map is the standard one (QMF bands 3..8 = bins 8..13, then [9,11) [11,14) [14,18) [18,23) [23,35) [35,64), checked
      // against borders_group when the ROM is installed), so every accumulator is a register.
      const int4 *row4 = reinterpret_cast<const int4 *>(row);
      i32 gsh[6];
#pragma unroll
      for (int g = 0; g < 6; g++) gsh[g] = rom[kPsRomGroupShift + g];
      u32 acc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int c = 0; c < 16; c++) {
        const int4 vr = row4[c], vi = row4[16 + c];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int k = 4 * c + j;
          if (k < 3) continue;
          const i32 xr = j == 0 ? vr.x : (j == 1 ? vr.y : (j == 2 ? vr.z : vr.w));
          const i32 xi = j == 0 ? vi.x : (j == 1 ? vi.y : (j == 2 ? vi.z : vi.w));
          const i32 mul = k < lsb ? mul_lo : (k < usb ? mul_hb : 1);
          const int shr = k < lsb ? shr_lo : (k < usb ? shr_hb : 0);
          const i32 r = (i32)((u32)xr * (u32)mul) >> shr, i = (i32)((u32)xi * (u32)mul) >> shr;
          const u32 t = min((u32)pw(r) + (u32)pw(i), 0x7fffffffu);
          if (k < 9) {
            w.pwt[s_][5 + k] = t;  // bins 8..13 = QMF bands 3..8
          } else {
            const int g = k < 11 ? 0 : (k < 14 ? 1 : (k < 18 ? 2 : (k < 23 ? 3 : (k < 35 ? 4 : 5))));
            if (k < usb_eff) acc[g] += t >> gsh[g];
          }
        }
      }
#pragma unroll
      for (int g = 0; g < 6; g++) w.pwt[s_][14 + g] = min(acc[g], 0x7fffffffu);
      // bins 0..7: hybrid sub-subbands — bins 0 / 1 pair (0, 5) / (4, 1), bins 2..7 one sub-subband each
      const i32 *hre = w.hybL[s_], *him = w.hybL[s_] + 10;
#pragma unroll
      for (int bin = 0; bin < 8; bin++) {
        const int s1 = bin == 0 ? 0 : (bin == 1 ? 4 : (int)borders[bin + 2]), s2 = bin == 0 ? 5 : 1;
        i32 pwr = add_sat(pw(hre[s1]), pw(him[s1]));
        if (bin < 2) pwr = add_sat(add_sat(pwr, pw(hre[s2])), pw(him[s2]));
        w.pwt[s_][bin] = (u32)pwr;
      }
    }
    __syncwarp();

    // ---- transient detection per parameter bin (ps_dec.c:482-620) for all 32 slots: a recursion over the slots per bin that
    //      nothing else feeds, so it runs here with its state in registers (lane = bin) instead of once per slot through shared
    //      memory; the transient ratio replaces the bin's power in pwt[slot][bin] (column 20 = 0: "no transient steering").
    //      The division only runs when some bin of the slot is in a transient.
    if (lane < 20) {
      i32 pk = peak[lane], nrg = peak[20 + lane], pdk = peak[40 + lane];
#pragma unroll 2
      for (int slot = 0; slot < 32; slot++) {
        const i32 pwr = (i32)w.pwt[slot][lane];
        i32 pv = shl32(pwr, 1);
        if (pv < 0) pv = 0;
        pk = lsl(mul32x16(pk, 0x620a), 1);
        if (pv > pk) pk = pv;
        pdk = add_sat(lsl(mul32x16(pdk, 0x6000), 1), sub_sat(pk, pv) >> 2);
        nrg = add_sat(lsl(mul32x16(nrg, 0x6000), 1), pv >> 2);
        const i32 pd = add_sat(pdk, pdk >> 1);
        i32 tr = 0x7fff;
        if (pd > nrg) tr = sext16(divide16_pos_lo(nrg, pd));
        w.pwt[slot][lane] = (u32)tr;
      }
      peak[lane] = pk;
      peak[20 + lane] = nrg;
      peak[40 + lane] = pdk;
    } else if (lane == 20) {
      for (int slot = 0; slot < 32; slot++) w.pwt[slot][20] = 0;
    }
    __syncwarp();

    // the left row of the next slot is fetched one slot ahead (its latency was the top stall of the slot loop)
    i32 nAr = mat[lane], nAi = mat[64 + lane], nBr = mat[32 + lane], nBi = mat[96 + lane];
    int next_border = prm[kPsPrmBorder];  // border of envelope `env`, kept in a register (prm lives in global memory)
    // The slot loop exists twice, with the rotation variant as a compile-time constant: selected at run time inside one loop
    // the compiler hoisted part of the other variant's arithmetic above the branch (40 wasted instructions per slot).
    auto slot_loop = [&](auto ab_tag) {
      constexpr bool AS_BUILT = decltype(ab_tag)::value;
#pragma unroll 1
    for (int slot = 0; slot < 32; slot++) {
      // ---- ixheaacd_init_rot_env at PS envelope borders (ps_dec.c:714-854): lane = stereo group ----
      if (env < 7 && slot == next_border) {
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
          D11r = m16_shl(inv_len, sext16(h11 - hv11));
          D12r = m16_shl(inv_len, sext16(h12 - hv12));
          D21r = m16_shl(inv_len, sext16(h21 - hv21));
          D22r = m16_shl(inv_len, sext16(h22 - hv22));
          H11r = hv11; H12r = hv12; H21r = hv21; H22r = hv22;
          hv11 = h11; hv12 = h12; hv21 = h21; hv22 = h22;
        }
        env++;
        next_border = env < 7 ? prm[kPsPrmBorder + env] : -1;
        __syncwarp();
      }

      // ---- left row (pre-shifted), this slot's hybrid sub-subbands ----
      i32 lAr, lAi, lBr, lBi;  // bands lane and lane + 32 of the left input
      {
        const i32 mA = slot < 6 ? mA_ov : mA_lb, mB = slot < 6 ? mB_ov : mB_lb;
        const int rA = slot < 6 ? rA_ov : rA_lb, rB = slot < 6 ? rB_ov : rB_lb;
        lAr = (i32)((u32)nAr * (u32)mA) >> rA;
        lAi = (i32)((u32)nAi * (u32)mA) >> rA;
        lBr = (i32)((u32)nBr * (u32)mB) >> rB;
        lBi = (i32)((u32)nBi * (u32)mB) >> rB;
        if (slot < 31) {
          const i32 *row = mat + 128 * (slot + 1);
          nAr = row[lane]; nAi = row[64 + lane]; nBr = row[32 + lane]; nBi = row[96 + lane];
        }
      }
      if (lane < 10) {
        w.hyb[lane] = w.hybL[slot][lane];
        w.hyb[16 + lane] = w.hybL[slot][10 + lane];
      }
      const i32 *lre = w.hyb, *lim = w.hyb + 16;
      i32 *rre = w.hyb + 32, *rim = w.hyb + 48;

      // ---- all-pass decorrelators (ps_dec.c:236-448) ----
      i32 apr = 0, api = 0;  // decorrelated signal of QMF band sb (lanes 10..29)
      const i32 qin_r = __shfl_sync(full, lAr, sb), qin_i = __shfl_sync(full, lAi, sb);  // left input of QMF band sb
      if (ap_act) {
        int16_t *dl = ap_dl + ap_dl_stride * d_idx;
        const i32 r0 = dl[0], i0 = dl[1];
        i32 rin = rotr(r0, i0, f0r, f0i), iin = roti(r0, i0, f0r, f0i);
        const i32 inr = hy ? lre[sb] : qin_r, ini = hy ? lim[sb] : qin_i;
        dl[0] = (int16_t)round16(inr);
        dl[1] = (int16_t)round16(ini);
#pragma unroll
        for (int m = 0; m < 3; m++) {
          const int di = m == 0 ? d_ser0 : (m == 1 ? d_ser1 : d_ser2);
          int16_t *q = ap_q + ap_q_stride * di + ap_m_stride * m;
          const i32 q0 = q[0], q1 = q[1];
          i32 rt = rotr(q0, q1, fmr[m], fmi[m]), it = roti(q0, q1, fmr[m], fmi[m]);
          rt = sext16(rt - m16_shl(rin, dec[m]));
          it = sext16(it - m16_shl(iin, dec[m]));
          q[0] = (int16_t)(rin + m16_shl(rt, dec[m]));
          q[1] = (int16_t)(iin + m16_shl(it, dec[m]));
          rin = rt;
          iin = it;
        }
        const i32 t = (i32)w.pwt[slot][trbin];
        apr = shl32(rin * t, 1);
        api = shl32(iin * t, 1);
        if (hy) { rre[sb] = apr; rim[sb] = api; }
      }
      // right input of bands lane / lane + 32 before the rotation: all-pass output (bands 3..22, from lane band + 7),
      // plain delays (ps_dec.c:596-645) or zero at and above usb
      i32 rAr = 0, rAi = 0, rBr = 0, rBi = 0;
      if (AS_BUILT) {
        // the as-built rotation zeroes both outputs of every QMF band below usb and the right input is zero above it, so the
        // delayed samples are never looked at: only the delay lines move on
        const int us = sext16(ps_usb), kB = lane + 32;
        if (lane >= b20 && lane < min(us, b21)) {
          int16_t *d = st + kPsStLd + 24 * d_long + 2 * (lane - b20);
          d[0] = (int16_t)round16(lAr);
          d[1] = (int16_t)round16(lAi);
        }
        if (kB >= b20 && kB < min(us, b22)) {
          int16_t *d = kB < b21 ? st + kPsStLd + 24 * d_long + 2 * (kB - b20) : st + kPsStSd + 2 * (kB - b21);
          d[0] = (int16_t)round16(lBr);
          d[1] = (int16_t)round16(lBi);
        }
      } else {
      rAr = __shfl_sync(full, apr, min(lane + 7, 29));
      rAi = __shfl_sync(full, api, min(lane + 7, 29));
      {
        const int us = sext16(ps_usb);
        if (lane < 3 || lane >= 23) { rAr = 0; rAi = 0; }
        if (lane >= b20 && lane < min(us, b21)) {
          int16_t *d = st + kPsStLd + 24 * d_long + 2 * (lane - b20);
          const i32 r = d[0], i = d[1], t = (i32)w.pwt[slot][18];
          d[0] = (int16_t)round16(lAr);
          d[1] = (int16_t)round16(lAi);
          rAr = shl32(r * t, 1);
          rAi = shl32(i * t, 1);
        }
        const int kB = lane + 32;
        if (kB >= b20 && kB < min(us, b21)) {
          int16_t *d = st + kPsStLd + 24 * d_long + 2 * (kB - b20);
          const i32 r = d[0], i = d[1], t = (i32)w.pwt[slot][18];
          d[0] = (int16_t)round16(lBr);
          d[1] = (int16_t)round16(lBi);
          rBr = shl32(r * t, 1);
          rBi = shl32(i * t, 1);
        } else if (kB >= b21 && kB < min(us, b22)) {
          int16_t *d = st + kPsStSd + 2 * (kB - b21);
          const i32 r = d[0], i = d[1], t = (i32)w.pwt[slot][19];
          d[0] = (int16_t)round16(lBr);
          d[1] = (int16_t)round16(lBi);
          rBr = shl32(r * t, 1);
          rBi = shl32(i * t, 1);
        }
        if (lane >= ps_usb) { rAr = 0; rAi = 0; }
        if (kB >= ps_usb) { rBr = 0; rBi = 0; }
      }
      }
      d_long = sext16(d_long + 1);
      if (d_long >= 14) d_long = 0;
      d_idx = d_idx + 1 >= 2 ? 0 : d_idx + 1;
      d_ser0 = d_ser0 + 1 >= rd0 ? 0 : d_ser0 + 1;
      d_ser1 = d_ser1 + 1 >= rd1 ? 0 : d_ser1 + 1;
      d_ser2 = d_ser2 + 1 >= rd2 ? 0 : d_ser2 + 1;

      // ---- rotation (ps_dec.c:856-991) ----
      H11r = sext16(H11r + D11r); H12r = sext16(H12r + D12r); H21r = sext16(H21r + D21r); H22r = sext16(H22r + D22r);
      __syncwarp();
      if (lane < 10) {  // hybrid sub-subband s = stereo group s
        const int s_ = lane;
        const i32 a = add_sat(mul32x16(lre[s_], H11r), mul32x16(rre[s_], H21r));
        const i32 b = add_sat(mul32x16(lim[s_], H11r), mul32x16(rim[s_], H21r));
        const i32 c = add_sat(mul32x16(lre[s_], H12r), mul32x16(rre[s_], H22r));
        const i32 d = add_sat(mul32x16(lim[s_], H12r), mul32x16(rim[s_], H22r));
        w.hyb[s_] = shl32(a, 2);
        w.hyb[16 + s_] = shl32(b, 2);
        w.hyb[32 + s_] = shl32(c, 2);
        w.hyb[48 + s_] = shl32(d, 2);
      }
      // QMF bands 3..usb-1.  as_built: the reference's x86-64 gcc build reads its type-punned coefficient copy as
      // zero (see apply_rot in oracle/src/ps.c), so both outputs are 0 there.
      i32 oLAr = lAr, oLAi = lAi, oLBr = lBr, oLBi = lBi, oRAr = rAr, oRAi = rAi, oRBr = rBr, oRBi = rBi;
      if (AS_BUILT) {
        if (lane >= 3 && lane < ps_usb) { oLAr = oLAi = oRAr = oRAi = 0; }
        if (lane + 32 < ps_usb) { oLBr = oLBi = oRBr = oRBi = 0; }
      } else {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int k = lane + 32 * h;
          int g = 10;
          while (g < 21 && k >= borders[g + 1]) g++;
          const i32 h0 = __shfl_sync(full, H11r, g), h1 = __shfl_sync(full, H12r, g);
          const i32 h2 = __shfl_sync(full, H21r, g), h3 = __shfl_sync(full, H22r, g);
          if (k >= 3 && k < ps_usb) {
            const i32 lr = h ? lBr : lAr, li = h ? lBi : lAi, rr = h ? rBr : rAr, ri = h ? rBi : rAi;
            const i32 a = shl32(add_sat(mul32x16(lr, h0), mul32x16(rr, h2)), 2);
            const i32 b = shl32(add_sat(mul32x16(li, h0), mul32x16(ri, h2)), 2);
            const i32 c = shl32(add_sat(mul32x16(lr, h1), mul32x16(rr, h3)), 2);
            const i32 d = shl32(add_sat(mul32x16(li, h1), mul32x16(ri, h3)), 2);
            if (h) { oLBr = a; oLBi = b; oRBr = c; oRBi = d; }
            else { oLAr = a; oLAi = b; oRAr = c; oRAi = d; }
          }
        }
      }
      __syncwarp();
      // fold the hybrid sub-subbands back into QMF bands 0..2 (resolutions 8 -> 6, 2, 2): lane = 4 * band + component
      i32 fold = 0;
      if (lane < 12) {
        const int s_ = lane >> 2, c = lane & 3, o = s_ == 0 ? 0 : 6 + 2 * (s_ - 1), cnt = s_ == 0 ? 6 : 2;
        fold = w.hyb[16 * c + o];
#pragma unroll
        for (int q = 1; q < 6; q++) fold = add_sat(fold, q < cnt ? w.hyb[16 * c + o + q] : 0);  // add_sat(x, 0) = x
      }
      {
        const int src = 4 * min(lane, 2);
        const i32 f0 = __shfl_sync(full, fold, src), f1 = __shfl_sync(full, fold, src + 1);
        const i32 f2 = __shfl_sync(full, fold, src + 2), f3 = __shfl_sync(full, fold, src + 3);
        if (lane < 3) { oLAr = f0; oLAi = f1; oRAr = f2; oRAi = f3; }
      }
      // ---- ixheaacd_shiftrountine on the left row, store both rows ----
      {
        auto cs = [&](i32 v) {
          if (common_shift < 0) return v >> min(-common_shift, 31);
          if (common_shift > 0) return shl32_sat(v, min(common_shift, 31));
          return v;
        };
        i32 *lrow = mat + 128 * slot, *rrow = right + 128 * slot;
        lrow[lane] = cs(oLAr);
        lrow[32 + lane] = cs(oLBr);
        lrow[64 + lane] = cs(oLAi);
        lrow[96 + lane] = cs(oLBi);
        rrow[lane] = oRAr;
        rrow[32 + lane] = oRBr;
        rrow[64 + lane] = oRAi;
        rrow[96 + lane] = oRBi;
      }
      __syncwarp();
    }
    };
    if (as_built)
      slot_loop(std::true_type{});
    else
      slot_loop(std::false_type{});
    if (lane < 22) {  // mixing-matrix state back to shared memory
      h11v[2 * lane] = (int16_t)hv11; h11v[2 * lane + 1] = (int16_t)hv12; h21v[2 * lane] = (int16_t)hv21; h21v[2 * lane + 1] = (int16_t)hv22;
      H11[2 * lane] = (int16_t)H11r; H11[2 * lane + 1] = (int16_t)H12r; H21[2 * lane] = (int16_t)H21r; H21[2 * lane + 1] = (int16_t)H22r;
      d11[2 * lane] = (int16_t)D11r; d11[2 * lane + 1] = (int16_t)D12r; d21[2 * lane] = (int16_t)D21r; d21[2 * lane + 1] = (int16_t)D22r;
    }
    __syncwarp();

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
  if (args.rot_nosat)
    ps_frame_kernel<true><<<(unsigned)grid, kPsWarps * 32, 0, stream>>>(args);
  else
    ps_frame_kernel<false><<<(unsigned)grid, kPsWarps * 32, 0, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace xb
