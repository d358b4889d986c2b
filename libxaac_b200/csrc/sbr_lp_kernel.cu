// sbr_lp_kernel.cu — the whole fixed-point LOW-POWER (real-valued) SBR stage as ONE fused kernel for sm_100a (B200).
//
// The reference runs this path for stereo HE-AACv1 (low_pow_flag = 1, decoder/ixheaacd_sbrdecoder.c:408-419).
// One warp owns one unit (one frame of one SBR channel).  The unit's whole QMF matrix (2 LPC rows + 6 overlap + 32
// current slots, 64 real bands) lives in shared memory for the frame, so HBM sees only the stage's compulsory traffic:
// 2 KB PCM16 in, 4 KB PCM16 out, the per-channel state once in and once out, and the side info.
// Replaces, bit-exactly, the fixed branch of
//   ixheaacd_sbr_dec                      decoder/ixheaacd_sbr_dec.c:662-1310            (low_pow_flag = 1)
// with its callees
//   ixheaacd_rescale_x_overlap            decoder/ixheaacd_sbrdec_lpfuncs.c:453-527
//   ixheaacd_cplx_anal_qmffilt            decoder/generic/ixheaacd_qmf_dec_generic.c:590-741  (LP: winadd + ixheaacd_dct3_32 :63-239)
//   ixheaacd_expsubbandsamples / ixheaacd_adjust_scale   decoder/ixheaacd_env_calc.c:1159-1207, 1099-1157 (real)
//   ixheaacd_low_pow_hf_generator         decoder/ixheaacd_lpp_tran.c:843-954  (+ covariance :271-372, filter1_lp :665-833,
//                                         filt_step3_lp :629-663, invfilt_level_emphasis sbrdec_lpfuncs.c:735-767)
//   ixheaacd_calc_sbrenvelope             decoder/ixheaacd_env_calc.c:692-1015 with the LP leaves: energy estimation
//                                         :1211-1380, alias reduction :78-227, conv_ergtoamplitudelp :423-448,
//                                         harm_idx_zerotwolp / onethreelp :1564-1757
//   ixheaacd_cplx_synt_qmffilt            decoder/ixheaacd_qmf_dec.c:811-1129  (LP: ixheaacd_inv_modulation_lp = dct2_64
//                                         qmf_dec.c:72-211, generic:241-257; ixheaacd_sbr_qmfsyn64_winadd generic:1508-1542)
//
// Lane mappings: lanes = bands for everything that walks rows (coalesced / conflict-free row segments); lanes = TIME
// SLOTS for the two DCTs (dct3_32 of the 32 analysis slots, dct2_64 of the 32 synthesis slots run as 32 independent
// serial transforms, one per lane, on rows with an odd word stride) — the transforms are tiny, so this keeps every
// lane busy without any cross-lane exchange; lanes = output samples in the polyphase windows.
// No tensor cores: no dense contraction exists here (the DCTs are 16/32-point radix-4 FFTs whose butterfly order and
// truncation points are part of the bit-exact contract).
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"
#include "env_common.cuh"

namespace xb {

constexpr int kLpWarps = 12;
#ifndef LP_GROUP
#define LP_GROUP 4
#endif
constexpr int kLpGroup = LP_GROUP;  // warps per phase-barrier group
constexpr int LS = 65;  // word stride of a matrix row in shared memory (odd: column walks are conflict-free)

struct LpTab {  // block-shared tables (image built on the host by sbr_lp_build_tables)
  int16_t qmf_c[1280];
  i32 dct23[66];  // dct23_tw << 16
  i32 post[18];   // post_fft_tbl << 16
  i32 w16[24];    // w_16 << 16
  i32 w32[60];    // w_32 << 16
  i32 periodic;   // 1: qmf_c[i] == qmf_c[i + 640] (the time-invariant window forms below need it)
  i32 pad;
};

struct LpWarpS {
  // envelope work arrays; during the analysis window the 791 words est .. pre hold the unit's time history instead
  int16_t est[2 * kMaxB], gain[2 * kMaxB], noise[2 * kMaxB], sine[2 * kMaxB], orig[2 * kMaxB];
  i32 line[kMaxB];
  // pre | x are contiguous rows of LS words.  x rows 0,1: LPC state rows; x rows 2..39: QMF matrix rows 0..37.
  // During the synthesis the 9 rows before matrix row 0 (pre + the dead LPC rows) hold the 9 old blocks of the state
  // ring and matrix row s itself receives the 128 WORD16 state samples of slot s: rows of `pre` = window history.
  i32 pre[7 * LS];
  i32 x[40 * LS];
  int16_t side[740];  // XAAC_SIDE_ENV [656] | XAAC_SIDE_HF [80] | apply
  int16_t st[kEnvStWords];
  int16_t ring[320];
  int16_t deg[64];
  int16_t fvec[64];
  int8_t sine_mapped[64];
  int8_t alias_red[128];
  int16_t sf[8], misc[16];
  int16_t limv[4 * 13];  // per limiter band: {max gain m, e | boost m, e, energy sum m, e}
};
static_assert(offsetof(LpWarpS, pre) == 4 * 336 && offsetof(LpWarpS, x) == offsetof(LpWarpS, pre) + 4 * 7 * LS,
              "est..pre..x must be contiguous");

struct LpBlockS {
  LpTab tab;
  EnvRomS rom;
  LpWarpS w[kLpWarps];
};

XB_DEV i32 hm_(i32 a, i32 b) { return __mulhi(a, (i32)((u32)b & 0xffff0000u)); }  // ops32.h:134
XB_DEV i32 abs_w(i32 a) { return a < 0 ? wneg(a) : a; }                            // ops32.h:271 (wraps)
XB_DEV i32 abs_s(i32 a) { return a == (i32)0x80000000 ? 0x7fffffff : (a < 0 ? -a : a); }
XB_DEV i32 x86shl(i32 v, int c) { return lsl(v, c & 31); }
XB_DEV i32 x86sar(i32 v, int c) { return v >> (c & 31); }

// basic_funcs.c:130-152
XB_DEV i32 fix_div_lp(i32 op1, i32 op2) {
  i32 q = 0;
  u32 num = (u32)abs_w(op1 >> 1), den = (u32)abs_w(op2 >> 1);
  if (num != 0) {
#pragma unroll 1
    for (int k = 15; k > 0; k--) {
      q = lsl(q, 1);
      num <<= 1;
      if (num >= den) {
        num -= den;
        q++;
      }
    }
  }
  return ((op1 ^ op2) < 0) ? -q : q;
}

// env_calc.c:1176-1187 (real): headroom of [s0,s1) x [b0,b1); m = matrix row 0
XB_DEV int lp_headroom(const i32 *m, int b0, int b1, int s0, int s1, int lane) {
  i32 mx = 1;
#pragma unroll 1
  for (int l = s0; l < s1; l++)
#pragma unroll 1
    for (int k = b0 + lane; k < b1; k += 32) mx |= abs_nrm(m[LS * l + k]);
  mx = (i32)__reduce_or_sync(0xffffffffu, (unsigned)mx);
  return pnorm32(mx);
}
// env_calc.c:1111-1129 (real)
XB_DEV void lp_adjust(i32 *m, int b0, int b1, int s0, int s1, int shift, int lane) {
  if (shift == 0) return;
  shift = max(-31, min(31, shift));
#pragma unroll 1
  for (int l = s0; l < s1; l++)
#pragma unroll 1
    for (int k = b0 + lane; k < b1; k += 32) {
      const i32 v = m[LS * l + k];
      m[LS * l + k] = shift > 0 ? lsl(v, shift) : (v >> -shift);
    }
}

// generic:1736-1829 — in-place radix-4 stage on interleaved complex x (one lane, rolled)
XB_DEV void radix4_lane(const i32 *w, i32 *x, int groups, int span) {
#pragma unroll 1
  for (int g = 0; g < groups; g++) {
#pragma unroll 1
    for (int i = 0; i < span; i++) {
      i32 *e0 = x + 2 * (g * 4 * span + i), *e1 = e0 + 2 * span, *e2 = e0 + 4 * span, *e3 = e0 + 6 * span;
      const i32 *tw = w + 6 * i;
      const i32 si1 = tw[0], co1 = tw[1], si2 = tw[2], co2 = tw[3], si3 = tw[4], co3 = tw[5];
      const i32 a0 = e0[0], a1 = e0[1], b0 = e1[0], b1 = e1[1], c0 = e2[0], c1 = e2[1], d0 = e3[0], d1 = e3[1];
      const i32 xh0 = add_sat(a0, c0), xl0 = sub_sat(a0, c0), xh20 = add_sat(b0, d0), xl20 = sub_sat(b0, d0);
      const i32 xh1 = add_sat(a1, c1), xl1 = sub_sat(a1, c1), xh21 = add_sat(b1, d1), xl21 = sub_sat(b1, d1);
      const i32 xt0 = sub_sat(xh0, xh20), yt0 = sub_sat(xh1, xh21);
      const i32 xt1 = add_sat(xl0, xl21), xt2 = sub_sat(xl0, xl21);
      const i32 yt2 = add_sat(xl1, xl20), yt1 = sub_sat(xl1, xl20);
      e0[0] = add_sat(xh0, xh20);
      e0[1] = add_sat(xh1, xh21);
      e3[0] = lsl(wadd(__mulhi(yt2, si3), __mulhi(xt2, co3)), 1);
      e3[1] = lsl(wsub(__mulhi(yt2, co3), __mulhi(xt2, si3)), 1);
      e2[0] = lsl(wadd(__mulhi(yt0, si2), __mulhi(xt0, co2)), 1);
      e2[1] = lsl(wsub(__mulhi(yt0, co2), __mulhi(xt0, si2)), 1);
      e1[0] = lsl(wadd(__mulhi(yt1, si1), __mulhi(xt1, co1)), 1);
      e1[1] = lsl(wsub(__mulhi(yt1, co1), __mulhi(xt1, si1)), 1);
    }
  }
}

// generic:63-239 — DCT-III of one slot, in place in the slot's 64-word row: in = row[0..63] (window-add outputs), result
// in row[0..31].  The two twiddle passes run in registers (static indices), the 16-point FFT in place in shared memory.
// dig_rev_table4_16 = {0, 16} (checked at ROM install).
XB_DEV void dct3_32_lane(const LpTab &t, i32 *row) {
  i32 o[32];
  o[0] = row[48] >> 7;
  o[1] = 0;
#pragma unroll
  for (int n = 1; n < 16; n++) {
    const i32 t0 = add_sat(row[48 + n] >> 7, row[48 - n] >> 7);
    const i32 t1 = sub_sat(row[16 + n] >> 7, row[16 - n] >> 7);
    const i32 re = t.dct23[4 * n], im = t.dct23[4 * n + 1];
    o[2 * n] = wadd(__mulhi(t0, re), __mulhi(t1, im));
    o[2 * n + 1] = wadd(wneg(__mulhi(t1, re)), __mulhi(t0, im));
  }
  {
    const i32 re = t.dct23[64], im = t.dct23[65];
    const i32 t1 = sub_sat(row[32] >> 7, row[0] >> 7), t0 = t1;
    const i32 u2 = wadd(__mulhi(t0, re), __mulhi(t1, im));
    const i32 u3 = wadd(wneg(__mulhi(t1, re)), __mulhi(t0, im));
    i32 u0 = o[0], u1 = o[1];
    const i32 a = wsub(wneg(u1), u3), b = wsub(u0, u2);
    u0 = wadd(wadd(u0, u2), a);
    u1 = wadd(wsub(u1, u3), b);
    o[0] = u0 >> 1;
    o[1] = u1 >> 1;
  }
#pragma unroll
  for (int n = 1; n <= 8; n++) {
    const bool last = n == 8;
    const i32 u0 = o[2 * n], u1 = o[2 * n + 1], u3 = o[33 - 2 * n], u2 = o[32 - 2 * n];
    i32 re = t.post[16 - 2 * n];
    if (last) re = (i32)((u32)(-(re >> 16)) << 16);  // (WORD16)(-*tr), wrapping
    const i32 im = t.post[2 * n];
    const i32 t0 = wsub(u0, u2), t2 = wadd(u1, u3);
    const i32 t1 = wadd(u0, u2) >> 1, t3 = wsub(u1, u3) >> 1;
    if (!last) {
      const i32 v4 = wadd(__mulhi(t0, re), __mulhi(t2, im));
      const i32 v5 = wadd(wneg(__mulhi(t2, re)), __mulhi(t0, im));
      o[2 * n] = wsub(t1, v4);
      o[2 * n + 1] = wadd(t3, v5);
      o[33 - 2 * n] = wadd(wneg(t3), v5);
      o[32 - 2 * n] = wadd(t1, v4);
    } else {
      const i32 v4 = wsub(__mulhi(t0, re), __mulhi(t2, im));
      const i32 v5 = wadd(__mulhi(t2, re), __mulhi(t0, im));
      o[16] = wadd(t1, v4);
      o[17] = wadd(t3, v5);
    }
  }
#pragma unroll
  for (int i = 0; i < 32; i++) row[i] = o[i];
  radix4_lane(t.w16, row, 1, 4);
  // generic:1831-1932 — final radix-4 (no twiddles) with digit-reversed scatter: row[0..31] -> row[32..63]
  i32 *in = row + 32;
#pragma unroll 1
  for (int k = 0; k < 2; k++) {
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
      const i32 *c = row + 16 * k + 8 * half;
      const int q = 4 * k + 2 * half;
      const i32 xh0 = add_sat(c[0], c[4]), xh1 = add_sat(c[1], c[5]);
      const i32 xl0 = sub_sat(c[0], c[4]), xl1 = sub_sat(c[1], c[5]);
      const i32 zh0 = add_sat(c[2], c[6]), zh1 = add_sat(c[3], c[7]);
      const i32 zl0 = sub_sat(c[2], c[6]), zl1 = sub_sat(c[3], c[7]);
      in[q] = add_sat(xh0, zh0);
      in[q + 1] = add_sat(xh1, zh1);
      in[8 + q] = add_sat(xl0, zl1);
      in[8 + q + 1] = sub_sat(xl1, zl0);
      in[16 + q] = sub_sat(xh0, zh0);
      in[16 + q + 1] = sub_sat(xh1, zh1);
      in[24 + q] = sub_sat(xl0, zl1);
      in[24 + q + 1] = add_sat(xl1, zl0);
    }
  }
  // generic:216-238 — output permutation row[32..63] -> row[0..31]
  row[0] = in[0];
  row[2] = in[1];
#pragma unroll 1
  for (int q = 0; q < 7; q++) {
    row[1 + 4 * q] = in[3 + 2 * q];
    row[3 + 4 * q] = in[2 + 2 * q];
    row[30 - 4 * q] = in[19 + 2 * q];
    row[28 - 4 * q] = in[18 + 2 * q];
  }
  row[29] = in[17];
  row[31] = in[16];
}

// qmf_dec.c:72-211 + generic:241-257 — DCT-II of one slot, in place: the slot's 64-word row is replaced by its 128 WORD16
// state samples fs[0..127].  The three passes that permute (pretwdct2, the digit-reversed radix-2 stage, the final
// scatter) go through registers with static indices; the FFT stages and twiddle passes run in place in shared memory.
// dig_rev_table2_32 = {0, 64, 16, 80} (checked at ROM install).
XB_DEV void dct2_64_lane(const LpTab &t, i32 *x) {
  i32 v[64];
#pragma unroll
  for (int n = 0; n < 32; n++) {  // pretwdct2
    v[n] = x[2 * n];
    v[63 - n] = x[2 * n + 1];
  }
#pragma unroll
  for (int i = 0; i < 64; i++) x[i] = v[i];
  radix4_lane(t.w32, x, 1, 8);
  radix4_lane(t.w32 + 48, x, 4, 2);
  // generic:1934-2015 — final radix-2 with digit-reversed scatter
#pragma unroll
  for (int i = 0; i < 64; i++) v[i] = x[i];
#pragma unroll
  for (int blk = 0; blk < 4; blk++) {
    const int h2 = (blk & 1) * 16 + (blk >> 1) * 4;  // dig_rev_table2_32[blk] >> 2
#pragma unroll
    for (int half = 0; half < 2; half++) {
      const int c = (blk >> 1) * 32 + (blk & 1) * 8 + 16 * half;
      const int q = h2 + 2 * half;
      x[q] = add_sat(v[c], v[c + 2]);
      x[q + 1] = add_sat(v[c + 1], v[c + 3]);
      x[32 + q] = sub_sat(v[c], v[c + 2]);
      x[32 + q + 1] = sub_sat(v[c + 1], v[c + 3]);
      x[8 + q] = add_sat(v[c + 4], v[c + 6]);
      x[8 + q + 1] = add_sat(v[c + 5], v[c + 7]);
      x[40 + q] = sub_sat(v[c + 4], v[c + 6]);
      x[40 + q + 1] = sub_sat(v[c + 5], v[c + 7]);
    }
  }
  // fftposttw, qmf_dec.c:107-159
  x[0] = lsl(x[0], 1);
  x[1] = lsl(x[1], 1);
#pragma unroll 1
  for (int k = 1; k <= 16; k++) {
    const i32 t0 = x[2 * k], t1o = x[2 * k + 1], t3o = x[65 - 2 * k], t2 = x[64 - 2 * k];
    const i32 in2 = sub_sat(t3o, t1o), in1 = add_sat(t3o, t1o);
    const i32 t1 = sub_sat(t0, t2), t3 = add_sat(t0, t2);
    const i32 re = t.post[k], im = t.post[16 - k];
    const i32 v1 = lsl(wsub(__mulhi(in1, re), __mulhi(t1, im)), 1);
    const i32 v2 = lsl(wadd(__mulhi(t1, re), __mulhi(in1, im)), 1);
    x[2 * k] = add_sat(t3, v1);
    x[2 * k + 1] = add_sat(in2, v2);
    x[65 - 2 * k] = sub_sat(v2, in2);
    x[64 - 2 * k] = sub_sat(t3, v1);
  }
  // posttwdct2, qmf_dec.c:161-211, pass 1 in place: x[0] = fs[32], x[1] = fs[0] = fs[64], x[2+2q] = r1_q, x[3+2q] = i1_q
  {
    const i32 ore = x[0], oim = x[1];
    const long long s = ((long long)ore + (long long)oim) >> 1;
    const i32 ore1 = s >= 0x7fffffffLL ? 0x7fffffff : (s <= -0x80000000LL ? (i32)0x80000000 : (i32)s);
    const i32 last = sub_sat(ore, oim);
    x[0] = round16(shl32(ore1, 4));
    x[1] = round16(shl32(__mulhi(last, t.dct23[64]), 4));
  }
#pragma unroll 1
  for (int q = 0; q < 31; q++) {
    const i32 ire = x[2 + 2 * q], iim = x[3 + 2 * q];
    const i32 re = t.dct23[2 + 2 * q], im = t.dct23[3 + 2 * q];
    const i32 ore = sub_sat(__mulhi(ire, re), __mulhi(iim, im));
    const i32 oim = add_sat(__mulhi(iim, re), __mulhi(ire, im));
    x[2 + 2 * q] = round16(shl32(ore, 4));
    x[3 + 2 * q] = round16(shl32(oim, 4));
  }
  // pass 2: scatter to fs[33+q] = fs[31-q] = r1_q, fs[95-q] = i1_q, fs[97+q] = -i1_q (saturating), fs[96] = 0 (generic:255)
#pragma unroll
  for (int i = 0; i < 64; i++) v[i] = x[i];
  auto FS = [&](int pidx) -> i32 {
    if (pidx == 0 || pidx == 64) return v[1];
    if (pidx < 32) return v[2 + 2 * (31 - pidx)];
    if (pidx == 32) return v[0];
    if (pidx < 64) return v[2 + 2 * (pidx - 33)];
    if (pidx < 96) return v[3 + 2 * (95 - pidx)];
    if (pidx == 96) return 0;
    return neg16(v[3 + 2 * (pidx - 97)]);
  };
#pragma unroll
  for (int j = 0; j < 64; j++) x[j] = (FS(2 * j) & 0xffff) | (i32)((u32)FS(2 * j + 1) << 16);
}

XB_DEV i32 mac_noise(i32 sig, i32 rnd, i32 nz) {  // ixheaac_mac16x16in32_shl_sat(sig, extract16h(rnd), nz)
  const i32 p = (rnd >> 16) * nz;
  return add_sat(sig, p == 0x40000000 ? 0x7fffffff : shl32(p, 1));
}
XB_DEV i32 gain_shift(i32 v, int shift) { return shift > 0 ? x86shl(v, shift) : x86sar(v, -shift); }

constexpr i32 kLpFactor = 0x010b0000 * 2;

// ---------------------------------------------------------------------------------------------------------------
// the low-power envelope adjuster (env_calc.c:692-1015, low_pow_flag = 1).  m = matrix row 0.  Returns the error flag.
// ---------------------------------------------------------------------------------------------------------------
XB_DEV int lp_envelope(LpWarpS &w, const EnvRomS &rom, const i32 *rand_ph, int lane) {
  const unsigned full = 0xffffffffu;
  i32 *mat = w.x + 2 * LS;
  const int16_t *prm = w.side;
  int16_t *st = w.st;
  int16_t *sf = w.sf;
  const int num_env = prm[kEnvNumEnv], trans_env = prm[kEnvTransientEnv];
  const int16_t *border = prm + kEnvBorderVec, *freq_res = prm + kEnvFreqRes, *nborder = prm + kEnvNoiseBorderVec;
  const int num_nf = prm[kEnvNumNfBands];
  const int sb_start = prm[kEnvSubBandStart], sb_end = prm[kEnvSubBandEnd];
  const int max_qmf = prm[kEnvMaxQmfSubband];
  const int max_qmf_prev = w.misc[kMiscMaxQmfPrev];
  const int num_sub_bands = sb_end - sb_start, skip = max_qmf - sb_start, bands = num_sub_bands - skip;
  const int16_t *fnoise = prm + kEnvFreqNoise;
  const int16_t *sf_arr = prm + kEnvSfArr;
  const int16_t *noise_floor = prm + kEnvNoiseFloor;
  int16_t *filt_me = st + kEnvStFiltMe, *filt_noise = st + kEnvStFiltNoise;
  int16_t *fme = filt_me + 2 * skip, *fno = filt_noise + skip;
  const int sf_hb_in = sf[kSfHb], sf_ov_hb_in = sf[kSfOvHb], sf_lb_in = sf[kSfLb];
  if (num_sub_bands > kMaxB || num_sub_bands < 0 || skip < 0 || max_qmf + num_sub_bands > 64) return 1;

#pragma unroll 1
  for (int i = lane; i < 2 * kMaxB; i += 32) w.est[i] = w.gain[i] = w.noise[i] = w.sine[i] = w.orig[i] = 0;
#pragma unroll 1
  for (int i = lane; i < 64; i += 32) w.sine_mapped[i] = 8;
#pragma unroll 1
  for (int i = lane; i < 128; i += 32) w.alias_red[i] = 0;
  __syncwarp();
  {  // ixheaacd_map_sineflags (sbrdec_lpfuncs.c:529-560)
    const int nhi = prm[kEnvNumSfHi];
    const int16_t *fhi = prm + kEnvFreqHi;
#pragma unroll 1
    for (int i = lane; i < nhi; i += 32) {
      const int pidx = nhi - 1 - i;
      const int old = st[kEnvStHarmPrev + pidx];
      const int add = prm[kEnvAddHarmonics + i];
      st[kEnvStHarmPrev + pidx] = (int16_t)(int8_t)add;
      if (add) {
        const int q = ((fhi[i + 1] + fhi[i]) - (fhi[0] << 1)) >> 1;
        w.sine_mapped[q & 63] = old ? 0 : (int8_t)trans_env;
      }
    }
  }
  int adj_e, final_e = 0;
  {  // env_calc.c:772-791
    const int first = (max_qmf_prev > max_qmf ? max_qmf_prev : max_qmf) - sb_start;
    int mx = 0;
#pragma unroll 1
    for (int i = max(first, 0) + lane; i < num_sub_bands; i += 32) mx = max(mx, (int)filt_noise[i]);
    mx = __reduce_max_sync(full, mx);
    adj_e = (st[kEnvStNoiseE] - norm32(mx)) - 16;
  }
  {  // :793-841
    int off = 0;
#pragma unroll 1
    for (int i = 0; i < num_env; i++) {
      const int n = prm[kEnvNumSfLo + freq_res[i]];
      int mx = 0;
#pragma unroll 1
      for (int j = lane; j < n; j += 32) mx = max(mx, sf_arr[off + j] & 0x3f);
      mx = __reduce_max_sync(full, mx);
      off += n;
      const int t = ((mx - 16) + 13) >> 1;
      if (border[i] < 16 && t > adj_e) adj_e = sext16(t);
      if (border[i + 1] > 16 && t > final_e) final_e = sext16(t);
    }
  }
  __syncwarp();

  int err = 0, m_off = 0, nf_idx = 0;
#pragma unroll 1
  for (int env = 0; env < num_env; env++) {
    const int start = 2 * border[env], end = 2 * border[env + 1], fr = freq_res[env];
    if (start >= 38 || end > 38 || nf_idx >= 2) { err = 1; break; }
    if (border[env] == nborder[nf_idx + 1]) { noise_floor += num_nf; nf_idx++; }
    const bool noise_absc = (env == trans_env) || (env == st[kEnvStTransPrev]);
    const int input_e = 15 - sf_hb_in;
    const int num_sfb = prm[kEnvNumSfLo + fr];
    const int16_t *ftab = prm + (fr ? kEnvFreqHi : kEnvFreqLo);

    // ---- energy estimation (real matrix) ----
    if (prm[kEnvInterpolFreq]) {  // env_calc.c:1211-1296, low_pow_flag = 1
      const i32 inv_width = rom.inv_int[end - start];
#pragma unroll 1
      for (int c = lane; c < sb_end - max_qmf; c += 32) {
        const int k = max_qmf + c;
        i32 max_val = 1;
#pragma unroll 1
        for (int l = start; l < end; l++) max_val = max(max_val, abs_nrm(mat[LS * l + k]));
        const int pre = pnorm32(max_val) - 3;
        int shift = 16 - pre;
        i32 accu = 0;
#pragma unroll 1
        for (int l = start; l < end; l++) {
          const i32 a = mat[LS * l + k];
          const i32 ta = sext16(shift > 0 ? (a >> shift) : lsl(a, -shift));
          accu = wadd(accu, ta * ta);
        }
        if (accu != 0) {
          shift = -pnorm32(accu);
          const i32 sum_m = sext16(shr32_dir_sat_limit(accu, 16 + shift));
          w.est[2 * c] = (int16_t)mult16_shl_sat_(sum_m, inv_width);
          shift = shift - (pre << 1) + 1;
          w.est[2 * c + 1] = (int16_t)((input_e << 1) + shift + 1);
        } else {
          w.est[2 * c] = w.est[2 * c + 1] = 0;
        }
      }
    } else {  // env_calc.c:1298-1380, low_pow_flag = 1
      int first_li = -1;
#pragma unroll 1
      for (int j = 0; j < num_sfb; j++)
        if (ftab[j] >= max_qmf) { first_li = ftab[j]; break; }
      const int top = ftab[num_sfb];
      const i32 inv_width = rom.inv_int[end - start];
#pragma unroll 1
      for (int k0 = (first_li < 0 ? top : first_li); k0 < top; k0 += 32) {
        const int k = k0 + lane;
        const bool act = k < top;
        int li = 0, ui = 0;
        i32 orv = 1;
        if (act) {
          int j = 0;
          while (ftab[j + 1] <= k) j++;
          li = ftab[j];
          ui = ftab[j + 1];
#pragma unroll 1
          for (int l = start; l < end; l++) orv |= abs_nrm(mat[LS * l + k]);
          w.line[k - first_li] = orv;
        }
        __syncwarp();
        int pre = 0;
        if (act) {
          i32 mx = 1;
#pragma unroll 1
          for (int kk = li; kk < ui; kk++) mx |= w.line[kk - first_li];
          pre = pnorm32(mx) - 4;
        }
        __syncwarp();
        if (act) {
          const int s = min(16 - pre, 31);
          i32 line = 0;
#pragma unroll 1
          for (int l = start; l < end; l++) {
            const i32 ta = sext16(shr32_dir(mat[LS * l + k], s));
            line = add_sat(line, ta * ta);
          }
          w.line[k - first_li] = shr32(line, 9);
        }
        __syncwarp();
        if (act) {
          i32 accumulate = 0;
#pragma unroll 1
          for (int kk = li; kk < ui; kk++) accumulate = add_sat(accumulate, w.line[kk - first_li]);
          const int shift = pnorm32(accumulate);
          i32 sum_m = sext16(shr32_dir_sat_limit(accumulate, 16 - shift));
          i32 sum_e = 0;
          if (sum_m != 0) {
            sum_m = mult16_shl_sat_(sum_m, inv_width);
            sum_m = mult16_shl_sat_(sum_m, rom.inv_int[ui - li]);
            sum_e = ((input_e << 1) + 11) - shift - (pre << 1);
          }
          w.est[2 * (k - first_li)] = (int16_t)sum_m;
          w.est[2 * (k - first_li) + 1] = (int16_t)sum_e;
        }
        __syncwarp();
      }
    }
    if (ftab[0] < sb_start) { err = 1; break; }
    __syncwarp();

    // ---- gains per band (env_calc.c:616-688) + the alias-reduction eligibility flags ----
    {
      const int f0 = ftab[0], top = ftab[num_sfb];
      const int c0 = max(max_qmf, f0);
#pragma unroll 1
      for (int k = f0 + lane; k < top; k += 32) {
        int j = 0;
        while (ftab[j + 1] <= k) j++;
        const int li = ftab[j], ui = ftab[j + 1];
        bool present = false;
#pragma unroll 1
        for (int kk = li; kk < ui; kk++) present |= (env >= w.sine_mapped[kk - f0]);
        w.alias_red[k - sb_start] = (int8_t)!present;
        if (k < max_qmf) continue;
        const int c = k - c0;
        const i32 v = sf_arr[m_off + j];
        const i32 ref_e = sext16((v & 0x3f) - 16), ref_m = sext16(v & 0xffc0);
        int nb = 0, ui_noise = fnoise[1];
#pragma unroll 1
        for (int kk = f0; kk <= k; kk++)
          if (kk >= ui_noise) {
            nb++;
            ui_noise = fnoise[nb + 1];
          }
        const i32 nm = sext16(noise_floor[nb] & 0xffc0), ne = sext16((noise_floor[nb] & 0x3f) - 38);
        w.orig[2 * c] = (int16_t)ref_m;
        w.orig[2 * c + 1] = (int16_t)ref_e;
        w.sine[2 * c] = w.sine[2 * c + 1] = 0;
        subbandgain(ref_m, nm, w.est[2 * c], w.est[2 * c + 1], ne, ref_e, present, env >= w.sine_mapped[skip + c],
                    noise_absc, &w.gain[2 * c], &w.noise[2 * c], &w.sine[2 * c], rom);
      }
    }
    m_off += num_sfb;
    __syncwarp();

    // ---- noise limiter (env_calc.c:229-421).  The (mantissa, exponent) sums round, so their order is part of the
    // contract: lane = limiter band for the two accumulations, lane = band for everything that is per band ----
    {
      const int16_t *lim = prm + kEnvLimTbl;
      const int nlf = min((int)prm[kEnvNumLfBands], 12);
      const i32 lg_m = rom.lim_gains[2 * prm[kEnvLimiterGains]], lg_e = rom.lim_gains[2 * prm[kEnvLimiterGains] + 1];
#pragma unroll 1
      for (int c = lane; c < nlf; c += 32) {
        const int b0 = lim[c] > skip ? lim[c] - skip : 0, b1 = lim[c + 1] > skip ? lim[c + 1] - skip : 0;
        i32 om = 0, oe = 0, em = 0, ee = 0;
#pragma unroll 1
        for (int k = b0; k < b1; k++) {
          acc_add(om, oe, w.orig[2 * k], w.orig[2 * k + 1]);
          acc_add(em, ee, w.est[2 * k], w.est[2 * k + 1]);
        }
        int nv = 16 - pnorm32(om);
        if (nv > 0) { om >>= nv; oe += nv; }
        nv = 16 - pnorm32(em);
        if (nv > 0) { em >>= nv; ee += nv; }
        const i32 sum_m = sext16(om), sum_e = sext16(oe);
        i32 mg_m;
        i32 mg_e = sext16(mant_div(sum_m, sext16(em), mg_m, rom) + (sum_e - sext16(ee)) + 1);
        const i32 mt = shl32(mg_m * lg_m, 1);
        mg_e = sext16(mg_e + lg_e);
        nv = norm32(mt);
        mg_e = sext16(mg_e - nv);
        mg_m = sext16(lsl(mt, nv) >> 16);
        if (mg_e >= 34) { mg_m = 0x3000; mg_e = 34; }
        w.limv[4 * c] = (int16_t)mg_m;
        w.limv[4 * c + 1] = (int16_t)mg_e;
        w.limv[4 * c + 2] = (int16_t)sum_m;
        w.limv[4 * c + 3] = (int16_t)sum_e;
      }
      __syncwarp();
      int myc[2] = {-1, -1};
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int k = lane + 32 * h;
        if (k < bands) {
#pragma unroll 1
          for (int cc = 0; cc < nlf; cc++) {
            const int b0 = lim[cc] > skip ? lim[cc] - skip : 0, b1 = lim[cc + 1] > skip ? lim[cc + 1] - skip : 0;
            if (k >= b0 && k < b1) { myc[h] = cc; break; }
          }
        }
        if (myc[h] >= 0) {
          const i32 mg_m = w.limv[4 * myc[h]], mg_e = w.limv[4 * myc[h] + 1];
          const i32 gm = w.gain[2 * k], ge = w.gain[2 * k + 1];
          if (ge > mg_e || (ge == mg_e && gm > mg_m)) {
            i32 na_m;
            i32 na_e = sext16(mant_div(mg_m, gm, na_m, rom));
            na_e = sext16(na_e + (mg_e - ge) + 1);
            w.noise[2 * k] = (int16_t)(shl32_dir_sat_limit(shl32((i32)w.noise[2 * k] * na_m, 1), na_e) >> 16);
            w.gain[2 * k] = (int16_t)mg_m;
            w.gain[2 * k + 1] = (int16_t)mg_e;
          }
        }
      }
      __syncwarp();
#pragma unroll 1
      for (int c = lane; c < nlf; c += 32) {
        const int b0 = lim[c] > skip ? lim[c] - skip : 0, b1 = lim[c + 1] > skip ? lim[c + 1] - skip : 0;
        if (b0 >= b1) continue;
        const i32 sum_m = w.limv[4 * c + 2], sum_e = w.limv[4 * c + 3];
        i32 am = 0, ae = 0;
#pragma unroll 1
        for (int k = b0; k < b1; k++) {
          acc_add(am, ae, ((i32)w.gain[2 * k] * w.est[2 * k]) >> 15, w.gain[2 * k + 1] + w.est[2 * k + 1]);
          const i32 sm = w.sine[2 * k];
          const bool use_s = sm != 0;
          if (use_s || !noise_absc) acc_add(am, ae, use_s ? sm : (i32)w.noise[2 * k], use_s ? (i32)w.sine[2 * k + 1] : (i32)w.noise[2 * k + 1]);
        }
        int nv = 16 - norm32(am);
        if (nv > 0) { am >>= nv; ae += nv; }
        i32 bg_m;
        i32 bg_e = sext16(mant_div(sum_m, sext16(am), bg_m, rom));
        bg_e = sext16(bg_e + (sum_e - sext16(ae)) + 1);
        if (bg_e > 2 || (bg_e == 2 && bg_m > 0x5061)) { bg_m = 0x5061; bg_e = 2; }
        w.limv[4 * c] = (int16_t)bg_m;
        w.limv[4 * c + 1] = (int16_t)bg_e;
      }
      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int k = lane + 32 * h;
        if (myc[h] >= 0) {
          const i32 bg_m = w.limv[4 * myc[h]], bg_e = w.limv[4 * myc[h] + 1];
          w.gain[2 * k] = (int16_t)mult16_shl_(w.gain[2 * k], bg_m);
          w.sine[2 * k] = (int16_t)mult16_shl_(w.sine[2 * k], bg_m);
          w.noise[2 * k] = (int16_t)mult16_shl_(w.noise[2 * k], bg_m);
          w.gain[2 * k + 1] = (int16_t)(w.gain[2 * k + 1] + bg_e);
          w.sine[2 * k + 1] = (int16_t)(w.sine[2 * k + 1] + bg_e);
          w.noise[2 * k + 1] = (int16_t)(w.noise[2 * k + 1] + bg_e);
        }
      }
    }
    __syncwarp();

    // ---- alias reduction (env_calc.c:78-227): group scan by lane 0, then one lane per group ----
    {
      const int16_t *deg = w.deg + sb_start;
      const int nsb = num_sub_bands;
      int ngrp = 0;
      {  // the grouping walk of env_calc.c:92-123 on two ballot masks, executed redundantly by every lane
        unsigned long long cond = 0, ared = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int k = lane + 32 * h;
          const bool r = k < nsb && w.alias_red[k] != 0;
          const bool c = r && k < nsb - 1 && deg[k + 1] != 0;
          cond |= (unsigned long long)__ballot_sync(full, c) << (32 * h);
          ared |= (unsigned long long)__ballot_sync(full, r) << (32 * h);
        }
        int grouping = 0, i = 0, last = 0;
#pragma unroll 1
        for (int k = 0; k < nsb - 1; k++) {
          const bool c = (cond >> k) & 1;
          int fv = -1;
          if (c) {
            if (!grouping) {
              fv = k;
              grouping = 1;
            } else if (last + 3 == k) {
              fv = k + 1;
              grouping = 0;
            }
          } else if (grouping) {
            grouping = 0;
            fv = ((ared >> k) & 1) ? k + 1 : k;
          }
          if (fv >= 0) {
            if (lane == 0) w.fvec[i] = (int16_t)fv;
            last = fv;
            i++;
          }
        }
        if (grouping) {
          if (lane == 0) w.fvec[i] = (int16_t)nsb;
          i++;
        }
        ngrp = i >> 1;
      }
      __syncwarp();
#pragma unroll 1
      for (int g = lane; g < ngrp; g += 32) {
        const int b0 = w.fvec[2 * g], b1 = w.fvec[2 * g + 1];
        // ixheaacd_avggain_calc with flag = 1 (env_calc.c:1454-1562): (orig, est) = (est, gain)
        i32 om = 0, oe = 0, em = 0, ee = 0;
#pragma unroll 1
        for (int k = b0; k < b1; k++) {
          i32 tm = w.est[2 * k], te = w.est[2 * k + 1];
          acc_add(om, oe, tm, te);
          tm = sext16((tm * (i32)w.gain[2 * k]) >> 16);
          te = sext16(te + w.gain[2 * k + 1] + 1);
          acc_add(em, ee, tm, te);
        }
        int nv = 16 - pnorm32(om);
        if (nv > 0) { om >>= nv; oe += nv; }
        nv = 16 - pnorm32(em);
        if (nv > 0) { em >>= nv; ee += nv; }
        const i32 amp_m = sext16(em), amp_e = sext16(ee), se_m = sext16(om), se_e = sext16(oe);
        i32 gg_m;
        const i32 gg_e = sext16(mant_div(amp_m, se_m, gg_m, rom) + (amp_e - se_e) + 1);
        i32 nm = 0, ne = 0;
#pragma unroll 1
        for (int k = b0; k < b1; k++) {
          i32 alpha = deg[k];
          if (k < nsb - 1 && deg[k + 1] > alpha) alpha = deg[k + 1];
          const i32 gain_m = alpha * gg_m;
          const i32 oma = sext16(0x7fff - alpha);
          i32 tg_m = w.gain[2 * k], tg_e = w.gain[2 * k + 1];
          tg_m = (oma * tg_m) >> 15;
          const i32 d = gg_e - tg_e;
          if (d >= 0) {
            tg_e = gg_e;
            tg_m = shr32(tg_m, d);
            tg_m = (gain_m >> 15) + tg_m;
          } else {
            tg_m = shr32(gain_m, 15 - d) + tg_m;
          }
          w.gain[2 * k] = (int16_t)tg_m;
          w.gain[2 * k + 1] = (int16_t)tg_e;
          const i32 tm = ((i32)((u32)tg_m * (u32)(i32)w.est[2 * k])) >> 16;  // :180, untruncated mantissa, wrapping product
          const i32 te = tg_e + w.est[2 * k + 1] + 1;
          const i32 dd = te - ne;
          if (dd >= 0) {
            nm = tm + shr32(nm, dd);
            ne = te;
          } else {
            nm = shr32(tm, -dd) + nm;
          }
        }
        nv = 16 - pnorm32(nm);
        if (nv > 0) { nm >>= nv; ne += nv; }
        i32 comp_m;
        i32 comp_e = sext16(mant_div(amp_m, sext16(nm), comp_m, rom));
        comp_e = sext16(comp_e + amp_e - sext16(ne) + 1 + 1);
#pragma unroll 1
        for (int k = b0; k < b1; k++) {
          w.gain[2 * k] = (int16_t)(((i32)w.gain[2 * k] * comp_m) >> 16);
          w.gain[2 * k + 1] = (int16_t)(w.gain[2 * k + 1] + comp_e);
        }
      }
    }
    __syncwarp();

    // ---- energies -> amplitudes (env_calc.c:423-448), start-up / exponent equalisation (:495-516, 1017-1058) ----
    int noise_e = sext16(start < 32 ? adj_e : final_e);
    const bool start_up = st[kEnvStStartUp] != 0;
    __syncwarp();
#pragma unroll 1
    for (int k = lane; k < bands; k += 32) {
      mant_exp_sqrt(&w.sine[2 * k], rom);
      mant_exp_sqrt(&w.gain[2 * k], rom);
      mant_exp_sqrt(&w.noise[2 * k], rom);
      int shift = (noise_e - w.noise[2 * k + 1]) - 4;
      if (shift > 0) w.noise[2 * k] = (int16_t)x86sar((i32)w.noise[2 * k], shift);
      else w.noise[2 * k] = (int16_t)x86shl((i32)w.noise[2 * k], -shift);
      shift = w.sine[2 * k + 1] - noise_e;
      if (shift > 0) w.sine[2 * k] = (int16_t)sat16(lsl((i32)w.sine[2 * k], min(sext16(shift), 15)));
      else w.sine[2 * k] = (int16_t)x86sar((i32)w.sine[2 * k], sext16(-shift));
      if (start_up) {
        fme[2 * k] = w.gain[2 * k];
        fme[2 * k + 1] = w.gain[2 * k + 1];
        fno[k] = w.noise[2 * k];
      } else {
        const i32 fe = fme[2 * k + 1], fm = fme[2 * k], diff = w.gain[2 * k + 1] - fe;
        if (diff >= 0) {
          fme[2 * k + 1] = w.gain[2 * k + 1];
          fme[2 * k] = (int16_t)(fm >> (diff & 31));
        } else {
          const int reserve = norm32(fm) - 16;
          if (diff + reserve >= 0) {
            fme[2 * k] = (int16_t)lsl(fm, -diff);
            fme[2 * k + 1] = (int16_t)(fe + diff);
          } else {
            fme[2 * k] = (int16_t)lsl(fm, reserve);
            fme[2 * k + 1] = (int16_t)(fe - reserve);
            const int shift2 = -(reserve + diff);
            w.gain[2 * k] = (int16_t)((i32)w.gain[2 * k] >> (shift2 & 31));
            w.gain[2 * k + 1] = (int16_t)(w.gain[2 * k + 1] + shift2);
          }
        }
      }
    }
    __syncwarp();
    if (start_up && lane == 0) {
      st[kEnvStStartUp] = 0;
      st[kEnvStNoiseE] = (int16_t)noise_e;
    }
    __syncwarp();

    // ---- time-slot adjustment, low power (env_calc.c:518-584, 1564-1757) ----
    int ph_index = st[kEnvStPhIndex], harm = st[kEnvStHarmIndex];
    int filt_noise_e = st[kEnvStNoiseE];
    const int nsb = num_sub_bands;
    const int n1 = nsb - 1;
    // everything that depends only on the band is hoisted out of the slot loop: lane owns bands lane and lane + 32
    i32 h_gm[2], h_ge[2], h_sm[2], h_nz[2], h_add[2];
    bool h_act[2], h_tone[2];
    {
      int tone_carry = 0;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int k = lane + 32 * h;
        h_act[h] = k < nsb;
        h_gm[h] = h_act[h] ? (i32)w.gain[2 * k] : 0;
        h_ge[h] = h_act[h] ? (i32)w.gain[2 * k + 1] : 0;
        h_sm[h] = h_act[h] ? (i32)w.sine[2 * k] : 0;
        h_nz[h] = h_act[h] ? (i32)w.noise[2 * k] : 0;
        const unsigned bal = __ballot_sync(full, h_act[h] && h_sm[h] != 0);
        h_tone[h] = (tone_carry + __popc(bal & (0xffffffffu >> (31 - lane)))) <= 16;
        tone_carry += __popc(bal);
        h_add[h] = 0;
        if (k >= 1 && k < n1) h_add[h] = mul32x16(kLpFactor, sext16((i32)w.sine[2 * (k - 1)] - (i32)w.sine[2 * (k + 1)]));
        else if (k == n1 && k >= 1) h_add[h] = mul32x16(kLpFactor, (i32)w.sine[2 * (k - 1)]);  // tms of the last band
        else if (k == 0) h_add[h] = mul32x16(kLpFactor, nsb > 1 ? (i32)w.sine[2] : 0);        // tm2 of the first band
      }
    }
    const int finv_base = (max_qmf & 1) ? -1 : 1;  // finv = !finv; finv = (finv << 1) - 1
#pragma unroll 1
    for (int l = start; l < end; l++) {
      int scale_change;
      if (l < 32) scale_change = adj_e - input_e;
      else {
        scale_change = final_e - input_e;
        if (l == 32 && start < 32) {
          const int diff = final_e - noise_e;
          noise_e = sext16(final_e);
          if (diff > 0) for (int k = lane; k < bands; k += 32) w.noise[2 * k] = (int16_t)(w.noise[2 * k] >> (diff & 31));
          else if (diff < 0) for (int k = lane; k < bands; k += 32) w.noise[2 * k] = (int16_t)lsl((i32)w.noise[2 * k], (-diff) & 31);
#pragma unroll
          for (int h = 0; h < 2; h++) h_nz[h] = h_act[h] ? (i32)w.noise[2 * (lane + 32 * h)] : 0;
        }
      }
      {
        const int diff = filt_noise_e - noise_e;
        if (diff > 0) for (int k = lane; k < num_sub_bands; k += 32) filt_noise[k] = (int16_t)(filt_noise[k] >> (diff & 31));
        else if (diff < 0) for (int k = lane; k < num_sub_bands; k += 32) filt_noise[k] = (int16_t)lsl((i32)filt_noise[k], (-diff) & 31);
        filt_noise_e = noise_e;
      }
      const int sc = scale_change - 1;
      i32 *re = mat + LS * l + max_qmf;
      const i32 *rnd = rand_ph + ph_index + 1;
      const bool odd = (harm & 1) != 0;
      const int finv0 = (harm == 3) ? -finv_base : finv_base;
      const int nz = sext16((noise_e - 16) - sext16(15 - sf_lb_in));
#pragma unroll
      for (int h = 0; h < 2; h++) {
        if (!h_act[h]) continue;
        const int k = lane + 32 * h;
        i32 s = gain_shift(mul32x16(re[k], h_gm[h]), h_ge[h] - sc);
        const i32 sk = h_sm[h];
        const bool noisy = sk == 0 && !noise_absc;
        if (noisy) s = mac_noise(s, __ldg(rnd + k), h_nz[h]);
        if (!odd) {  // ixheaacd_harm_idx_zerotwolp_dec (env_calc.c:1564-1615)
          if (!noisy) {
            const i32 sl = lsl(sk, 16);
            s = harm == 0 ? add_sat(s, sl) : sub_sat(s, sl);
          }
        } else if (k == 0) {  // ixheaacd_harm_idx_onethreelp (env_calc.c:1617-1757): first band
          i32 tm = mul32x16(kLpFactor, sk);
          if (nz > 0) tm = shl32(tm, nz);
          else tm = shr32(tm, -nz);
          if (finv0 < 0) {
            if (max_qmf > 0) re[-1] = add_sat(re[-1], tm);
            s = sub_sat(s, h_add[h]);
          } else {
            if (max_qmf > 0) re[-1] = sub_sat(re[-1], tm);
            s = add_sat(s, h_add[h]);
          }
        } else if (k < n1) {  // middle bands: finv alternates from band 1 on
          if (h_tone[h]) {
            const bool neg = ((k - 1) & 1) ? (finv0 > 0) : (finv0 < 0);
            s = add_sat(s, neg ? wneg(h_add[h]) : h_add[h]);
          }
        } else if (h_tone[h]) {  // last band (k == n1 >= 1)
          const bool plus = ((n1 - 1) & 1) ? (finv0 < 0) : (finv0 > 0);
          const i32 tm2 = mul32x16(kLpFactor, sk);
          if (plus) {
            s = add_sat(s, h_add[h]);
            if (k + max_qmf < 62) re[k + 1] = sub_sat(re[k + 1], tm2);
          } else {
            s = sub_sat(s, h_add[h]);
            if (k + max_qmf < 62) re[k + 1] = add_sat(re[k + 1], tm2);
          }
        }
        re[k] = s;
      }
      ph_index = (ph_index + nsb) & 511;
      harm = (harm + 1) & 3;
    }
    __syncwarp();
#pragma unroll 1
    for (int k = lane; k < bands; k += 32) {  // env_calc.c:1060-1078
      fme[2 * k] = w.gain[2 * k];
      fno[k] = w.noise[2 * k];
    }
    if (lane == 0) {
      st[kEnvStPhIndex] = (int16_t)ph_index;
      st[kEnvStHarmIndex] = (int16_t)harm;
      st[kEnvStNoiseE] = (int16_t)filt_noise_e;
    }
    __syncwarp();
  }
  if (err) return 1;

  {  // env_calc.c:956-1013
    const int first_start = border[0] * 2;
    int ov_reserve = 0, reserve = 0;
    __syncwarp();
    if (prm[kEnvChannelMode] == 3) {
      ov_reserve = lp_headroom(mat, max_qmf, sb_end, 0, first_start, lane);
      reserve = lp_headroom(mat, max_qmf, sb_end, first_start, 32, lane);
    }
    const int ov_adj_e = 15 - sf_ov_hb_in;
    const int output_e = max(ov_adj_e - ov_reserve, adj_e - reserve);
    lp_adjust(mat, max_qmf, sb_end, 0, first_start, ov_adj_e - output_e, lane);
    lp_adjust(mat, max_qmf, sb_end, first_start, prm[kEnvNumTimeSlots] * prm[kEnvTimeStep], adj_e - output_e, lane);
    if (lane == 0) {
      sf[kSfHb] = (int16_t)(15 - output_e);
      sf[kSfOvHb] = (int16_t)(15 - final_e);
      st[kEnvStTransPrev] = (trans_env == num_env) ? 0 : -1;
    }
  }
  __syncwarp();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// ixheaacd_low_pow_hf_generator (lpp_tran.c:843-954).  lane = low band.
// ---------------------------------------------------------------------------------------------------------------
XB_DEV void lp_hf_generator(LpWarpS &w, i32 *bw_prev_g, int norm_max, int lane) {
  const unsigned full = 0xffffffffu;
  i32 *x = w.x, *m = w.x + 2 * LS;
  const int16_t *env = w.side, *hf = w.side + kSideHf;
  const int num_patches = hf[0], num_columns = hf[3];
  const int16_t *patch = hf + 14, *bw_borders = hf + 4;
  const int num_if_bands = hf[kHfNumIfBands], max_qmf_subband = env[kEnvMaxQmfSubband];
  const int16_t *border = env + kEnvBorderVec;
  const int start_idx = sext16(border[0] * env[kEnvTimeStep]);
  const int stop_idx = num_columns + sext16(env[kEnvTimeStep] * sat16(border[env[kEnvNumEnv]] - env[kEnvNumTimeSlots]));
  const i32 nbw0 = 0x00000000, nbw06 = 0x4ccccccd, nbw075 = 0x60000000, nbw09 = 0x73333333, nbw098 = 0x7d70a3d7;

  // inverse-filter level emphasis (sbrdec_lpfuncs.c:735-767): lane i < num_if_bands owns bw_array[i]
  i32 my_bw = 0;
  if (lane < num_if_bands && lane < 6) {
    const int pm = w.misc[kMiscInvfPrev + lane], cm = hf[kHfInvf + lane];
    const i32 nb = (cm == 3) ? nbw098 : (cm == 2) ? nbw09 : (cm == 1) ? ((pm == 0) ? nbw06 : nbw075)
                                                                : ((pm == 1) ? nbw06 : nbw0);
    const i32 prev = bw_prev_g[lane];
    const i32 w1 = nb < prev ? 0x6000 : 0x7400, w2 = nb < prev ? 0x2000 : 0x0c00;
    i32 acc = wadd(lsl(mul32x16(nb, w1), 1), lsl(mul32x16(prev, w2), 1));
    if (acc < 0x02000000) acc = 0;
    if (acc >= 0x7f800000) acc = 0x7f800000;
    my_bw = acc;
    bw_prev_g[lane] = acc;
  }
  const int16_t *lastp = patch + 6 * (num_patches - 1);
  const int actual_stop = sext16(lastp[3] + lastp[5]);
  {  // :867-885
    const int len = min(6, stop_idx);
#pragma unroll 1
    for (int l = start_idx; l < len; l++)
#pragma unroll 1
      for (int b = actual_stop + lane; b < 64; b += 32) m[LS * l + b] = 0;
    if (actual_stop < 32) {
#pragma unroll 1
      for (int l = max(len, 0); l < stop_idx; l++)
        if (actual_stop + lane < 32) m[LS * l + actual_stop + lane] = 0;
    }
  }
  const int start_patch = max(1, (int)sext16(hf[1] - 2));
  const int stop_patch = min((int)patch[3], 32);
  __syncwarp();

  const int lb = start_patch + lane;
  const bool active = lb < stop_patch;
  // covariance of band lb over scratch rows 0..39 (lpp_tran.c:271-372)
  i32 c11 = 0, c22 = 0, c01 = 0, c02 = 0, c12 = 0, d = 0;
  if (active && norm_max != 30) {
    i32 x2 = shr32(x[lb], 3), x1 = shr32(x[LS + lb], 3);
    const i32 h10 = hm_(x1, x2), h00 = hm_(x2, x2);
    i32 p01 = 0, p02 = 0, p11 = 0, h3938 = 0, h3838 = 0;
#pragma unroll 2
    for (int n = 2; n < 40; n++) {
      const i32 x0 = shr32(x[LS * n + lb], 3);
      const i32 a = hm_(x0, x1), c = hm_(x1, x1);
      p01 = wadd(p01, a);
      p02 = wadd(p02, hm_(x0, x2));
      p11 = wadd(p11, c);
      if (n == 39) {
        h3938 = a;
        h3838 = c;
      }
      x2 = x1;
      x1 = x0;
    }
    const i32 p12 = wadd(wsub(p01, h3938), h10);
    const i32 p22 = wadd(wsub(p11, h3838), h00);
    const i32 mx = abs_nrm(p01) | abs_nrm(p02) | abs_nrm(p12) | p11 | p22;
    const int q = pnorm32(mx) & 31;
    c11 = lsl(p11, q); c22 = lsl(p22, q); c01 = lsl(p01, q); c02 = lsl(p02, q); c12 = lsl(p12, q);
    d = sub_sat(mul32(c22, c11), mul32(c12, c12));
  }
  // filter1_lp (lpp_tran.c:665-833): LPC coefficients and reflection coefficient of every low band
  i32 alpha0 = 0, alpha1 = 0, k1 = 0;
  if (active) {
    if (d != 0) {
      const int nd = norm32(d);
      const i32 inv = sext16(fix_div_lp(0x40000000, lsl(d, nd)));
      const i32 mod_d = abs_w(d);
      i32 t = sub_sat(mul32(c01, c12), mul32(c02, c11)) >> 2;
      if (abs_w(t) < mod_d) {
        const i32 v = (t == (i32)0x80000000 && inv == -32768) ? 0x7fffffff : lsl(mul32x16(t, inv), 1);
        alpha1 = sext16(lsl(v, nd) >> 15);
      }
      t = sub_sat(mul32(c02, c12), mul32(c01, c22)) >> 2;
      if (abs_w(t) < mod_d) {
        const i32 v = (t == (i32)0x80000000 && inv == -32768) ? 0x7fffffff : lsl(mul32x16(t, inv), 1);
        alpha0 = sext16(lsl(v, nd) >> 15);
      }
    }
    if (c11 == 0) k1 = 0;
    else if (abs_s(c01) >= c11) k1 = c01 < 0 ? 0x7fff : -0x8000;
    else k1 = sext16(-sext16(fix_div_lp(c01, c11)));
  }
  {  // alias degrees: the serial walk over lb becomes two shuffles (k1 of the two bands below)
    i32 k1b = __shfl_up_sync(full, k1, 1), k1b2 = __shfl_up_sync(full, k1, 2);
    if (lane < 1) k1b = 0;
    if (lane < 2) k1b2 = 0;
    i32 own = 0, wbv = 0;
    int wb = 0;
    if (active && lb > 1) {
      const i32 deg = sat16(0x7fff - mult16_shl_sat_(k1b, k1b));
      if (((lb & 1) == 0) && k1 < 0) {
        if (k1b < 0) {
          own = 0x7fff;
          if (k1b2 > 0) { wb = 1; wbv = deg; }
        } else if (k1b2 > 0) own = deg;
      }
      if (((lb & 1) != 0) && k1 > 0) {
        if (k1b > 0) {
          own = 0x7fff;
          if (k1b2 < 0) { wb = 1; wbv = deg; }
        } else if (k1b2 < 0) own = deg;
      }
    }
    const int nwb = __shfl_down_sync(full, wb, 1);
    const i32 nwbv = __shfl_down_sync(full, wbv, 1);
    if (active) w.deg[lb] = (int16_t)((lane < 31 && nwb) ? nwbv : own);
  }
  // patches: 2-tap real LPC FIR per (low band, patch); bw index depends only on the high band (monotone scan)
#pragma unroll 1
  for (int pt = 0; pt < num_patches; pt++) {
    const int16_t *pp = patch + 6 * pt;
    const int hb = lb + pp[4];
    const bool go = active && lb >= pp[0] && lb < pp[1] && hb >= max_qmf_subband && hb < 64;
    int idx = 0;
    while (idx < 5 && hb >= bw_borders[idx]) idx++;
    const i32 bw32 = __shfl_sync(full, my_bw, idx);
    if (!go) continue;
    i32 bw = sext16(bw32 >> 16);
    const i32 a0r = shl32(bw * alpha0, 1);
    bw = mult16_shl_sat_(bw, bw);
    const i32 a1r = shl32(bw * alpha1, 1);
    const int len = stop_idx - start_idx - 1;
    const i32 *lo = x + lb + LS * start_idx;
    i32 *hi = x + hb + LS * (start_idx + 2);
    if (len < 0) continue;
    if (bw > 0) {
      const int cnt = min(2 * (len / 2 + 1), 40 - (start_idx + 2));
      i32 p2 = lo[0], p1 = lo[LS];
#pragma unroll 2
      for (int t = 0; t < cnt; t++) {
        const i32 cur = lo[LS * (t + 2)];
        hi[LS * t] = add_sat(cur >> 2, lsl(wadd(hm_(p2, a1r), hm_(p1, a0r)), 1));
        p2 = p1;
        p1 = cur;
      }
    } else {
#pragma unroll 2
      for (int t = 0; t <= len; t++) hi[LS * t] = lo[LS * (t + 2)] >> 2;
    }
  }
  __syncwarp();
  // :927-951 — alias degrees follow the patches (destination bands lie above every source band)
#pragma unroll 1
  for (int base = hf[1]; base < hf[2]; base += 32) {
#pragma unroll 1
    for (int pt = 0; pt < num_patches; pt++) {
      const int16_t *pp = patch + 6 * pt;
      const int l2 = base + lane;
      const int hb = l2 + pp[4];
      int v = 0;
      const bool go = l2 < hf[2] && l2 >= pp[0] && l2 < pp[1] && hb < 64 && hb != pp[3] && l2 >= 0 && l2 < 64 && hb >= 0;
      if (go) v = w.deg[l2];
      __syncwarp();
      if (go) w.deg[hb] = (int16_t)v;
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(kLpWarps * 32, 1) sbr_dec_lp_kernel(SbrLpArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LpBlockS &sm = *reinterpret_cast<LpBlockS *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;
  {
    const i32 *src = reinterpret_cast<const i32 *>(p.lp_rom);
    i32 *dst = reinterpret_cast<i32 *>(&sm.tab);
    for (int i = threadIdx.x; i < (int)(sizeof(LpTab) / 4); i += blockDim.x) dst[i] = src[i];
    const int16_t *e = reinterpret_cast<const int16_t *>(p.env_rom);
    const int16_t *mr = reinterpret_cast<const int16_t *>(p.misc_rom);
    for (int i = threadIdx.x; i < 8; i += blockDim.x) sm.rom.lim_gains[i] = e[i];
    for (int i = threadIdx.x; i < 4; i += blockDim.x) sm.rom.smooth[i] = e[kERomSmooth / 2 + i];
    for (int i = threadIdx.x; i < 49; i += blockDim.x) sm.rom.inv_int[i] = e[kERomInvInt / 2 + i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sm.rom.inv_table[i] = mr[kMRomInvTable / 2 + i];
    for (int i = threadIdx.x; i < 257; i += blockDim.x) sm.rom.sqrt_table[i] = mr[kMRomSqrtTable / 2 + i];
  }
  __syncthreads();
  const LpTab &tab = sm.tab;
  const i32 *rand_ph = reinterpret_cast<const i32 *>(p.env_rom + kERomRandPh);
  LpWarpS &w = sm.w[warp];
  i32 *x = w.x, *m = w.x + 2 * LS;
  const long long warps_total = (long long)gridDim.x * kLpWarps;

#pragma unroll 1
  for (long long u0 = (long long)blockIdx.x * kLpWarps; u0 < p.n_units; u0 += warps_total) {
    const long long u = u0 + warp;
    // The warps of a block walk the phases of the stage together (block barrier between phases) so that they execute
    // the same few KB of this 120 KB kernel at the same time: without it the instruction cache thrashes
    // (profiles/r1_lp.md).  Only the last, partially filled iteration of a block runs unsynchronised.
    // Barrier groups of kLpGroup warps (named barriers) rather than the whole block: the groups drift apart, so one
    // group's global loads overlap another group's arithmetic.
    const int grp = warp / kLpGroup;
    const bool sync_ok = u0 + (grp + 1) * kLpGroup <= p.n_units;
#define LP_PHASE_SYNC()                                                                        \
  do {                                                                                         \
    if (sync_ok) asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(kLpGroup * 32) : "memory"); \
  } while (0)
    if (u >= p.n_units) continue;
    __syncwarp();
    {  // pull the next unit of this warp into L2 while this one is processed (its loads then cost an L2 hit, not DRAM)
      const long long un = u + warps_total;
      if (un < p.n_units) {
        auto pf = [&](const void *base, int bytes) {
          const char *q = reinterpret_cast<const char *>(base);
          for (int o = lane * 128; o < bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + o));
        };
        pf(p.side + un * kSideWords, 1480);
        if (p.w32) pf(p.w32 + un * 1024, 4096);
        else pf(p.time_in + un * p.in_unit_stride, p.in_ch == 1 ? 2048 : 0);
        pf(p.syn_states + un * 1280, 2560);
        pf(p.ov + un * 768, 1536);
        pf(p.anal_states + un * 320, 640);
        pf(p.env + un * kEnvStWords, 464);
        pf(p.lpc + un * 256, 640);
      }
    }
    // ---------------- load ----------------
    {
      // all of a lane's loads of the three records are issued before the first store (12 + 4 + 5 independent requests)
      const i32 *src = reinterpret_cast<const i32 *>(p.side + u * kSideWords);
      const i32 *ss = reinterpret_cast<const i32 *>(p.env + u * kEnvStWords);
      const i32 *rs = reinterpret_cast<const i32 *>(p.anal_states + u * 320);
      constexpr int nA = (370 + 31) / 32, nB = (kEnvStWords / 2 + 31) / 32, nC = 5;
      i32 va[nA], vb[nB], vc[nC];
#pragma unroll
      for (int q = 0; q < nA; q++) va[q] = lane + 32 * q < 370 ? __ldg(src + lane + 32 * q) : 0;  // words 0..739
#pragma unroll
      for (int q = 0; q < nB; q++) vb[q] = lane + 32 * q < kEnvStWords / 2 ? ss[lane + 32 * q] : 0;
#pragma unroll
      for (int q = 0; q < nC; q++) vc[q] = rs[lane + 32 * q];
#pragma unroll
      for (int q = 0; q < nA; q++)
        if (lane + 32 * q < 370) reinterpret_cast<i32 *>(w.side)[lane + 32 * q] = va[q];
#pragma unroll
      for (int q = 0; q < nB; q++)
        if (lane + 32 * q < kEnvStWords / 2) reinterpret_cast<i32 *>(w.st)[lane + 32 * q] = vb[q];
#pragma unroll
      for (int q = 0; q < nC; q++) reinterpret_cast<i32 *>(w.ring)[lane + 32 * q] = vc[q];
      if (lane < 8) w.sf[lane] = p.sf[u * 8 + lane];
      if (lane < 16) w.misc[lane] = p.misc[u * 16 + lane];
      const i32 *ov = p.ov + u * 768;
      {
        i32 vo[12];
#pragma unroll
        for (int l = 0; l < 6; l++) {
          vo[2 * l] = ov[64 * l + lane];
          vo[2 * l + 1] = ov[64 * l + 32 + lane];
        }
#pragma unroll
        for (int l = 0; l < 6; l++) {
          m[LS * l + lane] = vo[2 * l];
          m[LS * l + 32 + lane] = vo[2 * l + 1];
        }
      }
      const i32 *lpc = p.lpc + u * 256;
      x[lane] = lpc[lane];
      x[LS + lane] = lpc[128 + lane];
      x[32 + lane] = 0;
      x[LS + 32 + lane] = 0;
      w.deg[lane] = 0;
      w.deg[32 + lane] = 0;
    }
    __syncwarp();
    const int16_t *env = w.side;
    const int apply = w.side[kSideApply];
    const int16_t *border = env + kEnvBorderVec;
    const int num_env = env[kEnvNumEnv];
    if (lane == 0) w.sf[kSfLb] = 0;
    __syncwarp();

    // ---------------- ixheaacd_rescale_x_overlap (sbrdec_lpfuncs.c:453-527, real) ----------------
    if (apply) {
      const int old_lsb = w.misc[kMiscMaxQmfPrev], new_lsb = env[kEnvMaxQmfSubband];
      const int start_slot = env[kEnvTimeStep] * (w.misc[kMiscEndPosPrev] - env[kEnvNumTimeSlots]);
      const int syn_usb = w.misc[kMiscSynUsb];
      const int ov_lb = w.sf[kSfOvLb], ov_hb = w.sf[kSfOvHb];
      __syncwarp();
      if (lane == 0) {
        w.misc[kMiscCodecUsb] = (int16_t)new_lsb;
        w.misc[kMiscSynLsb] = (int16_t)new_lsb;
      }
      if (new_lsb != old_lsb && old_lsb > 0) {
        int b0 = min(old_lsb, new_lsb), b1 = max(old_lsb, new_lsb);
#pragma unroll 1
        for (int l = max(start_slot, 0); l < 6; l++)
#pragma unroll 1
          for (int k = old_lsb + lane; k < new_lsb; k += 32) m[LS * l + k] = 0;
        __syncwarp();
        int source_scale, target_scale, t_lsb, t_usb;
        if (new_lsb > old_lsb) { source_scale = ov_hb; target_scale = ov_lb; t_lsb = 0; t_usb = old_lsb; }
        else { source_scale = ov_lb; target_scale = ov_hb; t_lsb = old_lsb; t_usb = syn_usb; }
        const int ss = min(start_slot, 6);
        const int reserve = lp_headroom(m, b0, b1, 0, ss, lane);
        lp_adjust(m, b0, b1, 0, ss, reserve, lane);
        __syncwarp();
        source_scale += reserve;
        int delta = target_scale - source_scale;
        if (delta > 0) {
          delta = -delta;
          b0 = t_lsb;
          b1 = t_usb;
          if (lane == 0) w.sf[new_lsb > old_lsb ? kSfOvLb : kSfOvHb] = (int16_t)source_scale;
        }
        lp_adjust(m, b0, min(b1, 64), 0, ss, delta, lane);
      }
      __syncwarp();
    }

    LP_PHASE_SYNC();
    // ---------------- analysis: 5-tap window per slot (generic:528-588), lanes = outputs ----------------
    {
      const int16_t *pcm = p.time_in + u * p.in_unit_stride;
      const i32 *w32 = p.w32 ? p.w32 + u * 1024 : nullptr;
      const int qadj = p.w32 ? p.qshift_adj[u] : 0;
      auto samp = [&](int j) -> i32 {  // core-coder sample j of the frame as the WORD16 the reference's hand-over produces
        return w32 ? round16(shl32_sat(__ldg(w32 + j), qadj)) : (i32)pcm[(long long)p.in_ch * j];
      };
      int pos = p.anal_pos[2 * u], f1 = p.anal_pos[2 * u + 1], f2 = f1 + 64;
      // The reference keeps its 320-sample ring position and its coefficient phase in lock step (pos / 32 + filter_pos / 64
      // = 0 mod 10, both even, from reset on); then the window is the time-invariant FIR
      //   a[n] = sum_m x[t0 + 31 - n - 64 m] c[128 m + 2 n],  b[n] = sum_m x[t0 - 1 - n - 64 m] c[64 + 128 m + 2 n]
      // on the plain time history.  Any other (position, phase) pair runs the literal ring emulation below.
      const int P0 = pos >> 5, F0 = f1 >> 6;
      const bool lock = tab.periodic && (pos & 63) == 0 && (f1 & 127) == 0 && pos >= 0 && pos <= 256 && f1 >= 0 &&
                        f1 <= 512 && ((P0 + F0) % 10 == 0);
      if (lock) {
        int16_t *T = w.est;  // 1312 samples over est .. pre; T[j] = sample at time j - 288 relative to the frame start
#pragma unroll 1
        for (int j = lane; j < 288; j += 32) {
          int q = P0 + 9 - (j >> 5);
          if (q >= 10) q -= 10;
          T[j] = w.ring[32 * q + 31 - (j & 31)];
        }
        if (w32) {  // the whole frame's WORD32 samples, 2 x 16 requests in flight, converted on the way into the history
#pragma unroll
          for (int h = 0; h < 2; h++) {
            i32 vp[16];
#pragma unroll
            for (int q = 0; q < 16; q++) vp[q] = __ldg(w32 + 512 * h + lane + 32 * q);
#pragma unroll
            for (int q = 0; q < 16; q++) T[288 + 512 * h + lane + 32 * q] = (int16_t)round16(shl32_sat(vp[q], qadj));
          }
        } else if (p.in_ch == 1 && ((reinterpret_cast<uintptr_t>(pcm) & 3) == 0)) {
          const i32 *src = reinterpret_cast<const i32 *>(pcm);
          i32 *dst = reinterpret_cast<i32 *>(T + 288);
          i32 vp[16];  // the whole frame's PCM: 16 requests in flight
#pragma unroll
          for (int q = 0; q < 16; q++) vp[q] = __ldg(src + lane + 32 * q);
#pragma unroll
          for (int q = 0; q < 16; q++) dst[lane + 32 * q] = vp[q];
        } else {
#pragma unroll 4
          for (int j = lane; j < 1024; j += 32) T[288 + j] = (int16_t)samp(j);
        }
        i32 ca[5], cb[5];
#pragma unroll
        for (int mm = 0; mm < 5; mm++) {
          ca[mm] = tab.qmf_c[128 * mm + 2 * lane];
          cb[mm] = tab.qmf_c[64 + 128 * mm + 2 * lane];
        }
        __syncwarp();
        const int16_t *ta = T + 288 + 31 - lane, *tb = T + 288 - 1 - lane;
#pragma unroll 2
        for (int slot = 0; slot < 32; slot++) {
          i32 a = 0, b = 0;
#pragma unroll
          for (int mm = 0; mm < 5; mm++) {
            a += (i32)ta[32 * slot - 64 * mm] * ca[mm];
            b += (i32)tb[32 * slot - 64 * mm] * cb[mm];
          }
          m[LS * (6 + slot) + lane] = a;
          m[LS * (6 + slot) + 32 + lane] = b;
        }
        // ring after 32 slots: block q holds the latest slot i <= 31 with i = P0 - q (mod 10), newest sample first
#pragma unroll 1
        for (int pp = lane; pp < 320; pp += 32) {
          const int q = pp >> 5;
          int d = (31 - (P0 - q)) % 10;
          if (d < 0) d += 10;
          w.ring[pp] = T[288 + 32 * (31 - d) + 31 - (pp & 31)];
        }
        const int Pf = (P0 + 8) % 10;
        pos = 32 * Pf;
        f1 = 64 * ((10 - Pf) % 10);
      } else {
        i32 nxt = samp(lane);
#pragma unroll 1
        for (int slot = 0; slot < 32; slot++) {
          w.ring[pos + 31 - lane] = (int16_t)nxt;
          if (slot < 31) nxt = samp(32 * (slot + 1) + lane);
          __syncwarp();
          const int16_t *fp1 = w.ring + ((slot & 1) ? 32 : 0), *fp2 = w.ring + ((slot & 1) ? 0 : 32);
          i32 a = 0, b = 0;
#pragma unroll
          for (int j = 0; j < 5; j++) {
            a += (i32)fp1[lane + 64 * j] * (i32)tab.qmf_c[f1 + 2 * (lane + 64 * j)];
            b += (i32)fp2[lane + 64 * j] * (i32)tab.qmf_c[f2 + 2 * (lane + 64 * j)];
          }
          __syncwarp();
          pos -= 32;
          if (pos < 0) pos = 288;
          {
            const int n1 = f2 + 64, n2 = f1 + 64;
            f1 = n1;
            f2 = n2;
            if (f2 > 640) {
              f1 = 0;
              f2 = 64;
            }
          }
          m[LS * (6 + slot) + lane] = a;
          m[LS * (6 + slot) + 32 + lane] = b;
        }
      }
      __syncwarp();
      LP_PHASE_SYNC();
      // DCT-III of the 32 slots, lane = slot
      dct3_32_lane(tab, m + LS * (6 + lane));
      __syncwarp();
      if (lane == 0) {
        w.sf[kSfStLb] = 0;
        w.sf[kSfLb] = -10;  // generic:631
      }
      // ring + positions back to the state
      i32 *rd = reinterpret_cast<i32 *>(p.anal_states + u * 320);
#pragma unroll 1
      for (int i = lane; i < 160; i += 32) rd[i] = reinterpret_cast<const i32 *>(w.ring)[i];
      if (lane == 0) {
        p.anal_pos[2 * u] = (int16_t)pos;
        p.anal_pos[2 * u + 1] = (int16_t)f1;
      }
    }
    __syncwarp();

    LP_PHASE_SYNC();
    // ---------------- block floating point (sbr_dec.c:1050-1127, real) ----------------
    int save_lb_scale, max_samp_val;
    {
      const int usb = min((int)w.misc[kMiscCodecUsb], 32);
      int reserve = lp_headroom(m, 0, usb, 6, 38, lane);
      int reserve_ov1 = lp_headroom(m, 0, usb, 0, 6, lane);
      max_samp_val = min(reserve, reserve_ov1);
      const int reserve_ov2 = lp_headroom(x, 0, usb, 0, 2, lane);
      reserve_ov1 = min(reserve_ov1, reserve_ov2);
      const int lb0 = -10, ov_lb0 = w.sf[kSfOvLb];
      const int shift1 = lb0 + reserve, shift2 = ov_lb0 + reserve_ov1;
      const int min_shift = min(shift1, shift2);
      const int shift_over = shift2 - min_shift;
      reserve -= shift1 - min_shift;
      const int ov_shift = reserve_ov1 - shift_over;
      __syncwarp();
      lp_adjust(m, 0, usb, 0, 6, ov_shift, lane);
      lp_adjust(m, 0, usb, 6, 38, reserve, lane);
      lp_adjust(x, 0, usb, 0, 2, ov_shift, lane);
#pragma unroll 1
      for (int l = 6; l < 38; l++) m[LS * l + 32 + lane] = 0;  // ixheaacd_clr_subsamples
      save_lb_scale = lb0 + reserve;
      if (lane == 0) {
        w.sf[kSfOvLb] = (int16_t)(ov_lb0 + ov_shift);
        w.sf[kSfLb] = (int16_t)save_lb_scale;
      }
      save_lb_scale = sext16(save_lb_scale);
    }
    __syncwarp();

    int err = 0;
    if (apply) {
      lp_hf_generator(w, p.bw_prev + u * 6, max_samp_val, lane);
      if (lane == 0) w.sf[kSfHb] = (int16_t)(min((int)w.sf[kSfOvLb], (int)w.sf[kSfLb]) - 2);
      __syncwarp();
      err = lp_envelope(w, sm.rom, rand_ph, lane);
      err = __shfl_sync(full, err, 0);
      __syncwarp();
      if (!err) {
        const int16_t *hf = w.side + kSideHf;
        const int nif = hf[kHfNumIfBands];
        if (lane < nif && lane < 10) w.misc[kMiscInvfPrev + lane] = hf[kHfInvf + lane];
        if (lane == 0) {
          w.misc[kMiscMaxQmfPrev] = env[kEnvMaxQmfSubband];
          w.misc[kMiscEndPosPrev] = border[num_env];
        }
      }
    } else if (lane == 0) {
      w.sf[kSfHb] = (int16_t)save_lb_scale;
    }
    __syncwarp();

    LP_PHASE_SYNC();
    // ---------------- state that does not depend on the synthesis ----------------
    {
      i32 *lpc = p.lpc + u * 256;
      const int usb = min((int)w.misc[kMiscCodecUsb], 32);
      if (err) {  // the reference returns before the LPC / overlap update; its block-FP rescale of the LPC rows stays
        lpc[lane] = x[lane];
        lpc[128 + lane] = x[LS + lane];
      } else {
        lpc[lane] = lane < usb ? m[LS * 30 + lane] : x[lane];
        lpc[128 + lane] = lane < usb ? m[LS * 31 + lane] : x[LS + lane];
        i32 *ov = p.ov + u * 768;
#pragma unroll 1
        for (int l = 0; l < 6; l++) {
          ov[64 * l + lane] = m[LS * (32 + l) + lane];
          ov[64 * l + 32 + lane] = m[LS * (32 + l) + 32 + lane];
        }
      }
      i32 *sd = reinterpret_cast<i32 *>(p.env + u * kEnvStWords);
#pragma unroll 1
      for (int i = lane; i < kEnvStWords / 2; i += 32) sd[i] = reinterpret_cast<const i32 *>(w.st)[i];
      if (lane == 0 && p.err) p.err[u] = err ? (i32)0x80000000 : 0;
    }

    // ---------------- synthesis (qmf_dec.c:811-1129, low power) ----------------
    int off0 = p.syn_pos[2 * u], fpos0 = p.syn_pos[2 * u + 1];
    if (!err && ((off0 & 127) != 0 || (fpos0 & 63) != 0 || off0 < 0 || off0 >= 1280 || fpos0 < 0 || fpos0 >= 640)) {
      err = 2;  // ring positions the reference can never produce
      if (lane == 0 && p.err) p.err[u] = (i32)0x80000000;
    }
    i32 *const hist = w.pre;  // 9 old ring blocks in the 9 rows before matrix row 0; row 9 + s = matrix row s holds slot s
    const int b0 = off0 >> 7;
    if (!err) {
      const int st_syn = w.sf[kSfStSyn];
      const int sh_ov = max(-31, min(31, (st_syn - w.sf[kSfOvLb]) - 4));
      const int sh_lb = max(-31, min(31, (st_syn - w.sf[kSfLb]) - 4));
      const int sh_hb = max(-31, min(31, (st_syn - w.sf[kSfHb]) - 4));
      const int lsb = w.misc[kMiscSynLsb], usb = min((int)w.misc[kMiscSynUsb], 64);
#pragma unroll 1
      for (int l = 0; l < 32; l++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int k = lane + 32 * h;
          if (k < usb) {
            const int sh = k < lsb ? (l >= 6 ? sh_lb : sh_ov) : sh_hb;
            const i32 v = m[LS * l + k];
            m[LS * l + k] = sh > 0 ? lsl(v, sh) : (v >> -sh);
          }
        }
      }
      {
        const i32 *ss = reinterpret_cast<const i32 *>(p.syn_states + u * 1280);
        i32 vs[20];  // the synthesis ring (2.5 KB): 20 requests in flight, then the scatter into the history rows
#pragma unroll
        for (int q = 0; q < 20; q++) vs[q] = ss[lane + 32 * q];
#pragma unroll
        for (int q = 0; q < 20; q++) {
          const int i = lane + 32 * q, b = i >> 6;
          int a0 = b - b0;
          if (a0 < 0) a0 += 10;
          if (a0 != 0) hist[LS * (9 - a0) + (i & 63)] = vs[q];
        }
      }
    }
    __syncwarp();
    LP_PHASE_SYNC();
    // DCT-II of the 32 slots, lane = slot: matrix row -> 128 WORD16 state samples in history row 9 + slot
    if (!err) dct2_64_lane(tab, m + LS * lane);
    __syncwarp();
    LP_PHASE_SYNC();
    if (!err) {
      // 10-tap window (generic:1508-1542, shift = 2): lane -> outputs 2*lane, 2*lane+1 of every slot
      {
        int16_t *out = p.time_out + (u / p.out_ch) * (2048LL * p.out_ch) + (u % p.out_ch);
        auto store = [&](int i, i32 acc0, i32 acc1) {
          const i32 o0 = shl32_sat(acc0, 2) >> 16, o1 = shl32_sat(acc1, 2) >> 16;
          if (p.out_ch == 1) {
            *reinterpret_cast<i32 *>(out + 64 * i + 2 * lane) = (o0 & 0xffff) | (i32)((u32)o1 << 16);
          } else {
            out[(long long)p.out_ch * (64 * i + 2 * lane)] = (int16_t)o0;
            out[(long long)p.out_ch * (64 * i + 2 * lane + 1)] = (int16_t)o1;
          }
        };
        // Ring offset and coefficient phase are in lock step in the reference (drc_offset / 128 + filter_pos_syn / 64 = 0
        // mod 10, both even): the window is then the time-invariant FIR over the last 10 slots' state blocks,
        //   out_i[k] = sum_a block_{i-a}[64 (a & 1) + k] c[64 a + k]; any other pair runs the literal form below.
        const int f0s = fpos0 >> 6;
        const bool lock = tab.periodic && (off0 & 255) == 0 && (fpos0 & 127) == 0 && ((b0 + f0s) % 10 == 0);
        if (lock) {
          i32 clo[10], chi[10];
#pragma unroll
          for (int a = 0; a < 10; a++) {
            const i32 cv = *reinterpret_cast<const i32 *>(tab.qmf_c + 64 * a + 2 * lane);
            clo[a] = sext16(cv);
            chi[a] = cv >> 16;
          }
          const i32 *hp = hist + LS * 9 + lane;
#pragma unroll 2
          for (int i = 0; i < 32; i++) {
            i32 acc0 = 0x8000 >> 2, acc1 = 0x8000 >> 2;
#pragma unroll
            for (int a = 0; a < 10; a++) {
              const i32 hv = hp[LS * (i - a) + 32 * (a & 1)];
              acc0 += sext16(hv) * clo[a];
              acc1 += (hv >> 16) * chi[a];
            }
            store(i, acc0, acc1);
          }
        } else {
          int fpos = fpos0;
#pragma unroll 1
          for (int i = 0; i < 32; i++) {
            i32 acc0 = 0x8000 >> 2, acc1 = 0x8000 >> 2;
            int ab = (i - b0) % 10;  // age of ring block 0 at this slot
            if (ab < 0) ab += 10;
#pragma unroll 1
            for (int b = 0; b < 10; b++) {
              int a = ab + b;
              if (a >= 10) a -= 10;
              const int idx = ((i + b) & 1) * 32;  // word offset of the 64-sample phase inside the block
              const i32 hv = hist[LS * (9 + i - a) + idx + lane];
              const i32 cv = *reinterpret_cast<const i32 *>(tab.qmf_c + fpos + 64 * b + 2 * lane);
              acc0 += sext16(hv) * sext16(cv);
              acc1 += (hv >> 16) * (cv >> 16);
            }
            store(i, acc0, acc1);
            fpos += 64;
            if (fpos == 640) fpos = 0;
          }
        }
      }
      // ring after 32 slots: block b holds slot 31 - a, a = (b - b0 + 31) mod 10
      {
        i32 *sd = reinterpret_cast<i32 *>(p.syn_states + u * 1280);
#pragma unroll 1
        for (int i = lane; i < 640; i += 32) {
          const int b = i >> 6;
          int a = (b - b0 + 31) % 10;
          if (a < 0) a += 10;
          sd[i] = hist[LS * (40 - a) + (i & 63)];
        }
        if (lane == 0) {
          int off = (off0 - 128 * 32) % 1280;
          if (off < 0) off += 1280;
          p.syn_pos[2 * u] = (int16_t)off;
          p.syn_pos[2 * u + 1] = (int16_t)((fpos0 + 64 * 32) % 640);
          w.sf[kSfOvLb] = (int16_t)save_lb_scale;  // sbr_dec.c:1308
        }
      }
    }
    __syncwarp();
    if (lane < 8) p.sf[u * 8 + lane] = w.sf[lane];
    if (lane < 16) p.misc[u * 16 + lane] = w.misc[lane];
  }
}

size_t sbr_lp_table_bytes() { return sizeof(LpTab); }

// Builds the block-shared table image from the host's ia_qmf_dec_tables_struct prefix.  Returns 0, or -1 when the tables
// do not have the structure the kernel hard-codes (digit-reverse tables) or violate the no-saturation bound of the
// polyphase windows (sum of |coefficients| over the taps of one output).
int sbr_lp_build_tables(const uint8_t *qrom, uint8_t *out) {
  LpTab *t = reinterpret_cast<LpTab *>(out);
  const int16_t *c = reinterpret_cast<const int16_t *>(qrom + kQRomQmfC);
  const int16_t *d23 = reinterpret_cast<const int16_t *>(qrom + 772);
  const int16_t *pf = reinterpret_cast<const int16_t *>(qrom + 736);
  const int16_t *w16 = reinterpret_cast<const int16_t *>(qrom + kQRomW16);
  const int16_t *w32 = reinterpret_cast<const int16_t *>(qrom + kQRomW32);
  const int32_t *dr2 = reinterpret_cast<const int32_t *>(qrom + kQRomDigRev2_32);
  const int32_t *dr4 = reinterpret_cast<const int32_t *>(qrom + kQRomDigRev4_16);
  auto hi = [](int16_t v) { return (int32_t)((uint32_t)(uint16_t)v << 16); };
  if (dr2[0] != 0 || dr2[1] != 64 || dr2[2] != 16 || dr2[3] != 80 || dr4[0] != 0 || dr4[1] != 16) return -1;
  for (int i = 0; i < 1280; i++) t->qmf_c[i] = c[i];
  for (int i = 0; i < 66; i++) t->dct23[i] = hi(d23[i]);
  for (int i = 0; i < 18; i++) t->post[i] = hi(pf[i]);
  for (int i = 0; i < 24; i++) t->w16[i] = hi(w16[i]);
  for (int i = 0; i < 60; i++) t->w32[i] = hi(w32[i]);
  t->periodic = 1;
  for (int i = 0; i < 640; i++)
    if (c[i] != c[i + 640]) t->periodic = 0;
  t->pad = 0;
  for (int fpos = 0; fpos < 640; fpos += 64)
    for (int k = 0; k < 64; k++) {
      long long s = 0;
      for (int b = 0; b < 10; b++) s += c[fpos + 64 * b + k] < 0 ? -(long long)c[fpos + 64 * b + k] : c[fpos + 64 * b + k];
      if (s > 65535) return -1;
    }
  for (int base = 0; base <= 704; base += 64)
    for (int n = 0; n < 32; n++) {
      if (base + 2 * (n + 256) >= 1280) continue;
      long long s = 0;
      for (int j = 0; j < 5; j++) {
        const long long v = c[base + 2 * (n + 64 * j)];
        s += v < 0 ? -v : v;
      }
      if (s > 65535) return -1;
    }
  return 0;
}

cudaError_t launch_sbr_dec_lp(const SbrLpArgs &args, int num_sms, cudaStream_t stream) {
  static xb::PerDeviceOnce configured;
  const size_t smem = sizeof(LpBlockS);
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(sbr_dec_lp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.done();
  }
  long long need = (args.n_units + kLpWarps - 1) / kLpWarps;
  long long grid = num_sms;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  sbr_dec_lp_kernel<<<(unsigned)grid, kLpWarps * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace xb
