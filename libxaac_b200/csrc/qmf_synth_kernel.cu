// qmf_synth_kernel.cu — fixed-point complex ("HQ") 64-band SBR QMF synthesis for sm_100a (B200).
//
// One warp owns one unit (one frame of one output channel: 32 time slots x 64 complex bands -> 2048 PCM16
// samples) end to end; a persistent grid walks the batch.  Replaces, bit-exactly, the reference stage
//   ixheaacd_cplx_synt_qmffilt            decoder/ixheaacd_qmf_dec.c:811-1129        (non-PS, non-LP, non-ELD path)
// and the leaves it calls:
//   ixheaacd_adjust_scale_dec             decoder/ixheaacd_env_calc.c:1099           (block shifts, applied on load)
//   ixheaacd_inv_emodulation/cos_sin_mod  decoder/generic/ixheaacd_qmf_dec_generic.c:869, :259-466
//   ixheaacd_radix4bfly                   generic:1736-1829   (saturating radix-4, two stages of the 32-point FFT)
//   ixheaacd_postradixcompute2            generic:1934-2015   (final radix-2 + digit reversal)
//   ixheaacd_shiftrountine_with_rnd       generic:1638-1670   (fold to 128 WORD16 filter-state samples)
//   ixheaacd_sbr_qmfsyn64_winadd          generic:1508-1542   (10-tap polyphase window -> PCM16)
//
// Data flow per unit:
//   HBM matrix[32][128] WORD32 --LDG.32 coalesced, register double-buffered one slot pair ahead-->
//     block shift -> pre-twiddle -> smem T (two slots at a time so that the 2 x 16 radix-4 butterflies of a stage
//     fill all 32 lanes) -> radix-4, radix-4 -> radix-2 + post-twiddle + fold fused in registers
//     -> ring of 16 rows in shared memory, one per slot (sign-extended samples, one per word so a tap is one IMAD; output
//        pair k' owns words 4k'..4k'+3 of a row)
//     -> every four slots: 10-tap window over the 13 rows they need, coefficients by block age -> PCM16 pairs, STG.32 coalesced.
//   HBM filter_states[1280] WORD16 is read once and written once per unit in the reference's own (ring) layout.
// Algorithmic HBM bytes per unit: 16384 + 2560 + 2560 + 4096 = 25600.
//
// The saturating adds of the reference's window-add can never saturate with the standard prototype filter
// (sum of |coefficients| over the 10 taps <= 57308 < 65535, verified when the ROM is installed), so the taps are
// accumulated with wrapping multiply-adds, which is bit-identical.
//
// The same file holds the opt-in bulk-copy (TMA) staged variant qmf_synth_hq_tma_kernel (XAAC_B200_SYNTH_TMA=1).
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"
#include "qmf_synth_core.cuh"

namespace xb {

constexpr int kTHalf = 40;             // int2 per FFT half (32 used; +8 keeps the two halves on disjoint bank pairs)
constexpr int kTSlot = 2 * kTHalf;

// block-shared tables, pre-shifted on the host (qmf_synth_build_tables); the device image is this struct followed by qmf_c as
// 640 32-bit words (two consecutive WORD16 coefficients per word)
struct SynTables {
  int2 pre_tw[32];           // (wim<<16, wre<<16)  sbr_sin_cos_twiddle_l64
  int2 alt_tw[16];           // (wim<<16, wre<<16)  sbr_alt_sin_twiddle_l64
  int2 w1[24];               // radix-4 stage 1: position i -> (si,co) x 3, each << 16
  int2 w2[6];                // radix-4 stage 2: position i -> (si,co) x 3
  i32 postmap[32];           // F[p] = T[a] (+|-) T[a+1]: a | sign<<8
};
constexpr size_t kSynV2CoefOffset = sizeof(SynTables);

// slot-local swizzle of the FFT workspace (see DESIGN.md §QMF synthesis): keeps stage-1 (stride 8), stage-2
// (stride 2 inside groups of 8) and the pre-twiddle scatter conflict-free for 64-bit accesses.
XB_DEV int tsw(int e) { return e ^ (((e >> 3) & 3) << 1); }

// Saturating adds have no single-instruction form on sm_100 (add.sat.s32 lowers to IADD3 + 2 PLOP3 + 2 SEL), and the
// modulation is made of them.  Every slot pair is therefore classified first: if all its block-shifted inputs are
// below 2^fast_bits (a bound derived from the installed tables under which no add of the reference can saturate,
// see qmf_synth_build_tables), the pair runs with plain wrapping adds — bit-identical by construction; otherwise it
// runs the exact saturating code.  Real decoder data always takes the first path.
template <bool SAT> XB_DEV i32 A_(i32 a, i32 b) { return SAT ? add_sat(a, b) : wadd(a, b); }
template <bool SAT> XB_DEV i32 S_(i32 a, i32 b) { return SAT ? sub_sat(a, b) : wsub(a, b); }
template <bool SAT> XB_DEV i32 N_(i32 a) { return SAT ? neg_sat(a) : wneg(a); }

// Wrapping a+b / a-b as a 3-input add with a run-time zero: keeps the add on the ALU pipe (IADD3) instead of
// letting ptxas fold it into IMAD.HI's 64-bit addend, which costs extra register-pair moves on the (busier) FMA pipe.
XB_DEV i32 add3(i32 a, i32 b, i32 z) { return (i32)((u32)a + (u32)b + (u32)z); }
XB_DEV i32 sub3(i32 a, i32 b, i32 z) { return (i32)((u32)a - (u32)b + (u32)z); }

// one radix-4 butterfly, generic:1766-1822. e[m] = leg m (re,im); tw = 3 x (si<<16, co<<16)
template <bool SAT>
XB_DEV void radix4(int2 &e0, int2 &e1, int2 &e2, int2 &e3, const int2 t1, const int2 t2, const int2 t3, const i32 z) {
  i32 xh0 = A_<SAT>(e0.x, e2.x), xl0 = S_<SAT>(e0.x, e2.x);
  i32 xh20 = A_<SAT>(e1.x, e3.x), xl20 = S_<SAT>(e1.x, e3.x);
  i32 xh1 = A_<SAT>(e0.y, e2.y), xl1 = S_<SAT>(e0.y, e2.y);
  i32 xh21 = A_<SAT>(e1.y, e3.y), xl21 = S_<SAT>(e1.y, e3.y);
  i32 xt0 = S_<SAT>(xh0, xh20), yt0 = S_<SAT>(xh1, xh21);
  i32 xt1 = A_<SAT>(xl0, xl21), xt2 = S_<SAT>(xl0, xl21);
  i32 yt2 = A_<SAT>(xl1, xl20), yt1 = S_<SAT>(xl1, xl20);
  e0.x = A_<SAT>(xh0, xh20);
  e0.y = A_<SAT>(xh1, xh21);
  e3.x = lsl(add3(__mulhi(yt2, t3.x), __mulhi(xt2, t3.y), z), 1);
  e3.y = lsl(sub3(__mulhi(yt2, t3.y), __mulhi(xt2, t3.x), z), 1);
  e2.x = lsl(add3(__mulhi(yt0, t2.x), __mulhi(xt0, t2.y), z), 1);
  e2.y = lsl(sub3(__mulhi(yt0, t2.y), __mulhi(xt0, t2.x), z), 1);
  e1.x = lsl(add3(__mulhi(yt1, t1.x), __mulhi(xt1, t1.y), z), 1);
  e1.y = lsl(sub3(__mulhi(yt1, t1.y), __mulhi(xt1, t1.x), z), 1);
}

// rotation helpers for the (possibly saturating) pre/post twiddles
template <bool SAT> XB_DEV i32 RA_(i32 a, i32 b, i32 z) { return SAT ? add_sat(a, b) : add3(a, b, z); }
template <bool SAT> XB_DEV i32 RS_(i32 a, i32 b, i32 z) { return SAT ? sub_sat(a, b) : sub3(a, b, z); }

struct SynLane {       // per-lane constants of the modulation stages
  int s1base, s2base;  // FFT leg bases (leg m at base ^ swizzle handled via idx arrays below)
  int s1idx[4], s2idx[4];
  int pf_a, pf_b, pb_a, pb_b;  // radix-2 sources of the front / back complex of this lane
  i32 pf_sgn, pb_sgn;          // +1 / -1
  int pre_e;                   // pre-twiddle scatter slot
  int r16;
  int2 ptw, w1a, w1b, w1c, w2a, w2b, w2c, altb, altf;  // the lane's twiddles, loop-invariant
};

// Modulation of one slot pair: block-shifted inputs v[8] (a,b,c,d per slot; odd lanes hold them swapped, which turns
// the reference's alternating front/back pre-twiddle steps into one branch-free formula) -> fo[8] folded state samples
// (sign-extended WORD16 values) of the slot this lane serves (lanes 0-15: first slot, 16-31: second).
template <bool SAT, class SM>
XB_DEV void modulate_pair(const i32 *v, int2 *T, const SM &sm, const SynLane &L, int lane, i32 clamp_lo,
                          i32 clamp_hi, i32 fold_mul, i32 z, i32 *fo) {
  const int2 ptw = L.ptw;
  // ---- pre-twiddle (generic:290-367) ----
#pragma unroll
  for (int s = 0; s < 2; s++) {
    i32 a = v[4 * s + 0], b = v[4 * s + 1], c = v[4 * s + 2], d = v[4 * s + 3];
    int2 o1, o2;
    o1.x = RA_<SAT>(__mulhi(a, ptw.y), __mulhi(b, ptw.x), z);
    o1.y = RS_<SAT>(__mulhi(b, ptw.y), __mulhi(a, ptw.x), z);
    o2.x = RS_<SAT>(__mulhi(d, ptw.x), __mulhi(c, ptw.y), z);
    o2.y = RA_<SAT>(__mulhi(c, ptw.x), __mulhi(d, ptw.y), z);
    T[s * kTSlot + L.pre_e] = o1;
    T[s * kTSlot + kTHalf + L.pre_e] = o2;
  }
  __syncwarp();
  {  // ---- radix-4 stage 1 (span 8) ----
    int2 e0 = T[L.s1idx[0]], e1 = T[L.s1idx[1]], e2 = T[L.s1idx[2]], e3 = T[L.s1idx[3]];
    radix4<SAT>(e0, e1, e2, e3, L.w1a, L.w1b, L.w1c, z);
    T[L.s1idx[0]] = e0; T[L.s1idx[1]] = e1; T[L.s1idx[2]] = e2; T[L.s1idx[3]] = e3;
  }
  __syncwarp();
  {  // ---- radix-4 stage 2 (4 groups, span 2) ----
    int2 e0 = T[L.s2idx[0]], e1 = T[L.s2idx[1]], e2 = T[L.s2idx[2]], e3 = T[L.s2idx[3]];
    radix4<SAT>(e0, e1, e2, e3, L.w2a, L.w2b, L.w2c, z);
    T[L.s2idx[0]] = e0; T[L.s2idx[1]] = e1; T[L.s2idx[2]] = e2; T[L.s2idx[3]] = e3;
  }
  __syncwarp();
  // ---- radix-2 + digit reversal (generic:1934) + post-twiddle (generic:388-465) + fold (generic:1638) ----
  const int2 alt_b = L.altb, alt_f = L.altf;
  i32 G1[4], G2[4];  // [0]=G[2u] [1]=G[2u+1] [2]=G[62-2u] [3]=G[63-2u]
#pragma unroll
  for (int h = 0; h < 2; h++) {
    int2 fa = T[L.pf_a + h * kTHalf], fb = T[L.pf_b + h * kTHalf];
    int2 ba = T[L.pb_a + h * kTHalf], bb = T[L.pb_b + h * kTHalf];
    i32 Ff_r, Ff_i, Fb_r, Fb_i;
    if (SAT) {
      Ff_r = L.pf_sgn < 0 ? sub_sat(fa.x, fb.x) : add_sat(fa.x, fb.x);
      Ff_i = L.pf_sgn < 0 ? sub_sat(fa.y, fb.y) : add_sat(fa.y, fb.y);
      Fb_r = L.pb_sgn < 0 ? sub_sat(ba.x, bb.x) : add_sat(ba.x, bb.x);
      Fb_i = L.pb_sgn < 0 ? sub_sat(ba.y, bb.y) : add_sat(ba.y, bb.y);
    } else {
      Ff_r = (i32)((u32)fb.x * (u32)L.pf_sgn + (u32)fa.x);
      Ff_i = (i32)((u32)fb.y * (u32)L.pf_sgn + (u32)fa.y);
      Fb_r = (i32)((u32)bb.x * (u32)L.pb_sgn + (u32)ba.x);
      Fb_i = (i32)((u32)bb.y * (u32)L.pb_sgn + (u32)ba.y);
    }
    i32 *G = h ? G2 : G1;
    // front pair: words (2u, 2u+1) = (fim, fre) with alt[u-1]; u == 0 is the special first pair
    i32 fim = Ff_r, fre = Ff_i;
    i32 t_add = RA_<SAT>(__mulhi(fre, alt_f.x), __mulhi(fim, alt_f.y), z);
    i32 t_sub = h ? RS_<SAT>(__mulhi(fre, alt_f.y), __mulhi(fim, alt_f.x), z)
                  : RS_<SAT>(__mulhi(fim, alt_f.x), __mulhi(fre, alt_f.y), z);
    if (L.r16 == 0) {
      G[0] = h ? (Ff_i >> 1) : (Ff_r >> 1);
      G[3] = h ? N_<SAT>(Ff_r >> 1) : N_<SAT>(Ff_i >> 1);
    } else {
      G[0] = h ? t_sub : t_add;
      G[3] = h ? N_<SAT>(t_add) : t_sub;
    }
    // back pair: words (62-2u, 63-2u) = (im, re) with alt[u]
    i32 im = Fb_r, re = Fb_i;
    i32 b_add = RA_<SAT>(__mulhi(re, alt_b.y), __mulhi(im, alt_b.x), z);
    i32 b_sub = h ? RS_<SAT>(__mulhi(re, alt_b.x), __mulhi(im, alt_b.y), z)
                  : RS_<SAT>(__mulhi(im, alt_b.y), __mulhi(re, alt_b.x), z);
    G[2] = h ? b_sub : b_add;
    G[1] = h ? N_<SAT>(b_add) : b_sub;
  }
  auto R = [&](i32 x) {
    x = max(clamp_lo, min(clamp_hi, x));
    return (i32)((u32)x * (u32)fold_mul + 0x8000u) >> 16;
  };
  // j = 2u: r1=G1[0] i1=G2[0] r2=G1[3] i2=G2[3];  j = 2u+1: r1=G1[1] i1=G2[1] r2=G1[2] i2=G2[2]
  fo[0] = R(S_<SAT>(G2[0], G1[0]));  // st[2u]
  fo[1] = R(S_<SAT>(G2[1], G1[1]));  // st[2u+1]
  fo[2] = R(S_<SAT>(G2[2], G1[2]));  // st[62-2u]
  fo[3] = R(S_<SAT>(G2[3], G1[3]));  // st[63-2u]
  fo[4] = R(A_<SAT>(G2[3], G1[3]));  // st[64+2u]
  fo[5] = R(A_<SAT>(G2[2], G1[2]));  // st[65+2u]
  fo[6] = R(A_<SAT>(G2[1], G1[1]));  // st[126-2u]
  fo[7] = R(A_<SAT>(G2[0], G1[0]));  // st[127-2u]
}

// per-lane constants of the modulation stages (lanes 0-15 serve the first slot of a pair, 16-31 the second)
XB_DEV void syn_lane_setup(SynLane &L, int lane, const i32 *postmap) {
  const int fs_slot = lane >> 4;
  L.r16 = lane & 15;
  const int h1 = L.r16 >> 3, i1 = L.r16 & 7, g2 = (L.r16 >> 1) & 3, i2 = L.r16 & 1;
#pragma unroll
  for (int m = 0; m < 4; m++) {
    L.s1idx[m] = fs_slot * kTSlot + h1 * kTHalf + tsw(i1 + 8 * m);          // legs at i1 + 8m
    L.s2idx[m] = fs_slot * kTSlot + h1 * kTHalf + tsw(8 * g2 + i2 + 2 * m);  // legs at 8 g2 + i2 + 2m
  }
  const i32 pm_f = postmap[L.r16], pm_b = postmap[31 - L.r16];
  L.pf_a = fs_slot * kTSlot + tsw(pm_f & 255);
  L.pf_b = fs_slot * kTSlot + tsw((pm_f & 255) + 1);
  L.pb_a = fs_slot * kTSlot + tsw(pm_b & 255);
  L.pb_b = fs_slot * kTSlot + tsw((pm_b & 255) + 1);
  L.pf_sgn = ((pm_f >> 8) & 1) ? -1 : 1;
  L.pb_sgn = ((pm_b >> 8) & 1) ? -1 : 1;
  // pre-twiddle: step n = lane; even n -> complex n/2, odd n -> complex 31-(n-1)/2
  L.pre_e = tsw((lane & 1) ? 31 - (lane >> 1) : (lane >> 1));
}

// =====================================================================================================================
// The kernel: slot-pair modulation (above) + linear-time window, four slots per pass.
//
// Filter state: instead of the reference's ring with a rotating coefficient pointer (round 1 kept it, tap-major transposed: one
// window pass per slot = 5 LDS.128 + 10 LDS.64 + 20 IMAD and two warp syncs; 1.69 ms per 131 072 units) the folded blocks go to a
// ring of 16 ROWS in time order (row = slot & 15, 128 words: output pair k' owns words
// 4k'..4k'+3 = both 64-sample halves) and the window runs once per FOUR slots: the 13 rows s0-9 .. s0+3 are read once (13 LDS.128)
// for all 80 taps, and the coefficient of a block depends on its age only (see qmf_synth_core.cuh), so a tap is
// row[s - a][half(a)] * w_a with w_a read once per pass.  Shared memory per warp: 8 KB rows + 1.25 KB FFT workspace; one block of
// 16 warps per SM — four per scheduler, which leaves room for the lane's nine twiddle pairs in registers (a warp count that is
// not a multiple of four costs 10 %: the schedulers run unevenly loaded).
// =====================================================================================================================
constexpr int kG4Warps = 16;  // 4 per scheduler; measured: 12 -> 1.51 ms, 16 -> 1.37, 20 -> 1.38, 18 / 22 / 23 (uneven) -> 1.45-1.51

struct SynG4Warp {
  i32 rows[16 * 128];
  int2 T[2 * kTSlot];
};
struct SynG4Block {
  int2 pre_tw[32];
  int2 alt_tw[16];
  int2 w1[24];
  int2 w2[6];
  i32 postmap[32];
  int2 cw[20 * 32];  // cw[32 b + k'] = (qmf_c[64 (b mod 10) + 2k'], qmf_c[.. + 1]), sign-extended: tap of age a at block kappa + a
  SynG4Warp w[kG4Warps];
};
static_assert(sizeof(SynTables) == offsetof(SynG4Block, cw), "the block starts with the table image");

// window of slots s0 .. s0+3 (s0 = 4 g, GPH = g & 3 fixes the ring rows at compile time) for output pair `lane`
template <int P, int GPH>
XB_DEV void window_ring4(const i32 *rows, const int2 *cwk, int lane, i32 (&o)[4][2]) {
  i32 acc[4][2];
#pragma unroll
  for (int t = 0; t < 4; t++) acc[t][0] = acc[t][1] = 0x4000;
#pragma unroll
  for (int r = -9; r < 4; r++) {
    const int4 x = *reinterpret_cast<const int4 *>(rows + ((4 * GPH + r) & 15) * 128 + 4 * lane);
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int a = t - r;
      if (a < 0 || a > 9) continue;
      const int2 w = cwk[32 * a];
      const bool h = ((P + a) & 1) != 0;
      acc[t][0] += (h ? x.z : x.x) * w.x;
      acc[t][1] += (h ? x.w : x.y) * w.y;
    }
  }
#pragma unroll
  for (int t = 0; t < 4; t++) {
    o[t][0] = max(-0x40000000, min(0x3fffffff, acc[t][0])) >> 15;
    o[t][1] = max(-0x40000000, min(0x3fffffff, acc[t][1])) >> 15;
  }
}
template <int P>
XB_DEV void window_ring4_dispatch(int gph, const i32 *rows, const int2 *cwk, int lane, i32 (&o)[4][2]) {
  switch (gph) {
    case 0: window_ring4<P, 0>(rows, cwk, lane, o); break;
    case 1: window_ring4<P, 1>(rows, cwk, lane, o); break;
    case 2: window_ring4<P, 2>(rows, cwk, lane, o); break;
    default: window_ring4<P, 3>(rows, cwk, lane, o); break;
  }
}

__global__ void __launch_bounds__(kG4Warps * 32, 1)
qmf_synth_hq_kernel(QmfSynthArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SynG4Block &sm = *reinterpret_cast<SynG4Block *>(smem_raw);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  {
    const i32 *src = reinterpret_cast<const i32 *>(p.rom);
    i32 *dst = reinterpret_cast<i32 *>(&sm);
    for (int i = threadIdx.x; i < (int)(sizeof(SynTables) / 4); i += blockDim.x) dst[i] = src[i];
    const i32 *c32 = reinterpret_cast<const i32 *>(p.rom + kSynV2CoefOffset);
    for (int i = threadIdx.x; i < 20 * 32; i += blockDim.x) {
      const i32 v = c32[32 * ((i >> 5) % 10) + (i & 31)];
      sm.cw[i] = make_int2((i32)(int16_t)v, v >> 16);
    }
  }
  __syncthreads();
  i32 *rows = sm.w[warp].rows;
  int2 *T = sm.w[warp].T;
  const int warps_total = gridDim.x * kG4Warps;
  SynLane L;
  const int fs_slot = lane >> 4;
  syn_lane_setup(L, lane, sm.postmap);
  {
    const int i1 = L.r16 & 7, i2 = L.r16 & 1;
    L.ptw = sm.pre_tw[lane];
    L.w1a = sm.w1[3 * i1]; L.w1b = sm.w1[3 * i1 + 1]; L.w1c = sm.w1[3 * i1 + 2];
    L.w2a = sm.w2[3 * i2]; L.w2b = sm.w2[3 * i2 + 1]; L.w2c = sm.w2[3 * i2 + 2];
    L.altb = sm.alt_tw[L.r16]; L.altf = sm.alt_tw[L.r16 > 0 ? L.r16 - 1 : 0];
  }
  const int bandA = (lane & 1) ? 63 - lane : lane;
  const int bandB = 63 - bandA;

  for (long long u = (long long)blockIdx.x * kG4Warps + warp; u < p.n_units; u += warps_total) {
    if (p.gate && p.gate[u] == 0) continue;
    const i32 *mat = p.matrix + u * p.mat_stride;
    const int16_t *prm = p.params + u * 8;
    const int ov_lb_scale = prm[0], lb_scale = prm[1], hb_scale = prm[2], st_syn = prm[3];
    const int lsb = prm[4], usb = prm[5], split = prm[6];
    const int off0 = p.pos[2 * u], fpos0 = p.pos[2 * u + 1];
    const int Bw0 = off0 >> 7;
    int kappa = (fpos0 >> 6) + Bw0;
    if (kappa >= 10) kappa -= 10;
    const int2 *cwk = sm.cw + 32 * kappa + lane;
    int ov_lb_shift = (st_syn - ov_lb_scale) - 8, lb_shift = (st_syn - lb_scale) - 8;
    int hb_shift = (st_syn - hb_scale) - 8;
    const int out_shift = -(st_syn - 3) + 1;
    auto enc = [](int sh, i32 &mul, int &shr) {
      sh = max(-31, min(31, sh));
      mul = sh > 0 ? (i32)(1u << sh) : 1;
      shr = sh < 0 ? -sh : 0;
    };
    i32 mulA_ov, mulA_lb, mulB_ov, mulB_lb;
    int shrA_ov, shrA_lb, shrB_ov, shrB_lb;
    enc(bandA < lsb ? ov_lb_shift : (bandA < usb ? hb_shift : 0), mulA_ov, shrA_ov);
    enc(bandA < lsb ? lb_shift : (bandA < usb ? hb_shift : 0), mulA_lb, shrA_lb);
    enc(bandB < lsb ? ov_lb_shift : (bandB < usb ? hb_shift : 0), mulB_ov, shrB_ov);
    enc(bandB < lsb ? lb_shift : (bandB < usb ? hb_shift : 0), mulB_lb, shrB_lb);
    const bool noshift = __all_sync(0xffffffffu, (mulA_ov & mulA_lb & mulB_ov & mulB_lb) == 1 && (shrA_ov | shrA_lb | shrB_ov | shrB_lb) == 0);
    const i32 fast_lim = (i32)(1u << p.fast_bits);
    const i32 clamp_lo = (i32)0x80000000 >> out_shift;
    const i32 clamp_hi = (i32)(0x7fff7fffu >> out_shift);
    const i32 fold_mul = (i32)(1u << out_shift);

    // ---- history: ring block B of filter_states (reference layout) has age a = (B - Bw0) mod 10 -> row (-a) & 15 ----
    i32 *st32 = reinterpret_cast<i32 *>(p.states + u * 1280);
    {
      i32 wv[20];
#pragma unroll
      for (int t = 0; t < 20; t++) wv[t] = st32[32 * t + lane];
#pragma unroll
      for (int B = 0; B < 10; B++) {
        int a = B - Bw0;
        if (a < 0) a += 10;
        if (a != 0)
          *reinterpret_cast<int4 *>(rows + ((16 - a) & 15) * 128 + 4 * lane) =
              make_int4((i32)(int16_t)wv[2 * B], wv[2 * B] >> 16, (i32)(int16_t)wv[2 * B + 1], wv[2 * B + 1] >> 16);
      }
    }
    __syncwarp();

    i32 nx[8];
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const i32 *m = mat + 128 * s;
      nx[4 * s + 0] = __ldg(m + bandA);
      nx[4 * s + 1] = __ldg(m + bandB);
      nx[4 * s + 2] = __ldg(m + 64 + bandA);
      nx[4 * s + 3] = __ldg(m + 64 + bandB);
    }
    int16_t *pcm = p.pcm + (p.pcm_unit_stride ? u * p.pcm_unit_stride
                                               : ((p.ch_fac == 1) ? u * 2048 : (u / p.ch_fac) * (2048LL * p.ch_fac) + (u % p.ch_fac)));

#pragma unroll 1
    for (int pr = 0; pr < 16; pr++) {
      i32 v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = nx[j];
      if (!noshift) {  // the left channel after PS arrives in the synthesis scale already: no block shifts at all
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const bool ov = (2 * pr + s) < split;
          const i32 mA = ov ? mulA_ov : mulA_lb, mB = ov ? mulB_ov : mulB_lb;
          const int rA = ov ? shrA_ov : shrA_lb, rB = ov ? shrB_ov : shrB_lb;
          v[4 * s + 0] = (i32)((u32)v[4 * s + 0] * (u32)mA) >> rA;
          v[4 * s + 1] = (i32)((u32)v[4 * s + 1] * (u32)mB) >> rB;
          v[4 * s + 2] = (i32)((u32)v[4 * s + 2] * (u32)mA) >> rA;
          v[4 * s + 3] = (i32)((u32)v[4 * s + 3] * (u32)mB) >> rB;
        }
      }
      // all inputs inside [-2^fast_bits, 2^fast_bits): three-input min / max, one instruction per two values
      i32 vmx = max(max(v[0], v[1]), max(v[2], v[3])), vmn = min(min(v[0], v[1]), min(v[2], v[3]));
      vmx = max(vmx, max(max(v[4], v[5]), max(v[6], v[7])));
      vmn = min(vmn, min(min(v[4], v[5]), min(v[6], v[7])));
      const bool big = vmx >= fast_lim || vmn < -fast_lim;
      if (pr < 15) {
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const i32 *m = mat + 128 * (2 * pr + 2 + s);
          nx[4 * s + 0] = __ldg(m + bandA);
          nx[4 * s + 1] = __ldg(m + bandB);
          nx[4 * s + 2] = __ldg(m + 64 + bandA);
          nx[4 * s + 3] = __ldg(m + 64 + bandB);
        }
      }
      i32 fo[8];
      if (!__any_sync(0xffffffffu, big))
        modulate_pair<false>(v, T, sm, L, lane, clamp_lo, clamp_hi, fold_mul, p.zero, fo);
      else
        modulate_pair<true>(v, T, sm, L, lane, clamp_lo, clamp_hi, fold_mul, 0, fo);
      // ---- the fold of this lane's slot into its row: pairs r16 and 31 - r16, both halves ----
      {
        i32 *row = rows + ((2 * pr + fs_slot) & 15) * 128;
        *reinterpret_cast<int4 *>(row + 4 * L.r16) = make_int4(fo[0], fo[1], fo[4], fo[5]);
        *reinterpret_cast<int4 *>(row + 4 * (31 - L.r16)) = make_int4(fo[2], fo[3], fo[6], fo[7]);
      }
      if (pr & 1) {  // ---- four new rows are complete: window of slots 2 pr - 2 .. 2 pr + 1 ----
        __syncwarp();
        const int s0 = 2 * pr - 2;
        i32 o[4][2];
        if (Bw0 & 1)
          window_ring4_dispatch<1>((s0 >> 2) & 3, rows, cwk, lane, o);
        else
          window_ring4_dispatch<0>((s0 >> 2) & 3, rows, cwk, lane, o);
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const int slot = s0 + t;
          if (p.ch_fac == 1) {
            *reinterpret_cast<i32 *>(pcm + 64 * slot + 2 * lane) = (o[t][0] & 0xffff) | (i32)((u32)o[t][1] << 16);
          } else {
            pcm[p.ch_fac * (64 * slot + 2 * lane)] = (int16_t)o[t][0];
            pcm[p.ch_fac * (64 * slot + 2 * lane + 1)] = (int16_t)o[t][1];
          }
        }
        __syncwarp();
      }
    }

    // ---- the last ten blocks back to the reference's ring: slot r sits at ring block (Bw0 - r) mod 10 ----
    {
      int B = Bw0 + 8;  // slot 22
      if (B >= 10) B -= 10;
#pragma unroll
      for (int r = 22; r < 32; r++) {
        const int4 x = *reinterpret_cast<const int4 *>(rows + (r & 15) * 128 + 4 * lane);
        st32[32 * (2 * B) + lane] = (i32)(((u32)x.x & 0xffffu) | ((u32)x.y << 16));
        st32[32 * (2 * B + 1) + lane] = (i32)(((u32)x.z & 0xffffu) | ((u32)x.w << 16));
        B = B ? B - 1 : 9;
      }
      if (lane == 0) {
        int off = off0 + 1024, fpos = fpos0 + 128;  // 32 steps of -128 mod 1280 / +64 mod 640
        if (off >= 1280) off -= 1280;
        if (fpos >= 640) fpos -= 640;
        p.pos[2 * u] = (int16_t)off;
        p.pos[2 * u + 1] = (int16_t)fpos;
      }
    }
    __syncwarp();
  }
}

// =====================================================================================================================
// Variant (opt-in, XAAC_B200_SYNTH_TMA=1): lane = slot modulation in registers, matrix rows staged by 1-D bulk copies (TMA), linear-time window.
// See qmf_synth_core.cuh for the arithmetic and the layout; this part is the data movement.
//
// Per warp in shared memory: 41 rows x 132 words (9 rows of history + the 32 rows of the frame; the frame rows arrive by
// one 512-byte cp.async.bulk per lane, complete on the warp's mbarrier, and are overwritten in place by the folded
// filter-state samples) + the per-band block-shift table of the unit (2 variants x 64 bands x (mul, shr)).
// 8 warps per SM (22.7 KB each, 255 registers): a half FFT of a slot (32 complex words) lives in registers.
// =====================================================================================================================
using namespace syn;

constexpr int kSyn2Warps = 8;  // 2 per scheduler: 255 registers, no spills (10 warps at 168 registers measured slower)

struct Syn2Warp {
  i32 rows[kRows * kRowW];  // row -9 .. row 31
  int2 sh[2][64];           // [0]: slots < split (overlap scale), [1]: the others; per band (mul, shr)
};
struct Syn2Smem {
  Syn2Warp w[kSyn2Warps];
  unsigned long long bar[kSyn2Warps];
};

XB_DEV u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
XB_DEV void mbar_init(u32 bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
XB_DEV void mbar_arrive_expect_tx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
XB_DEV void bulk_g2s(u32 dst, const void *src, u32 bytes, u32 bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
XB_DEV void bulk_prefetch_l2(const void *src, u32 bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
XB_DEV void mbar_wait(u32 bar, u32 parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "XB_SYN_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra XB_SYN_DONE;\n"
      "bra XB_SYN_WAIT;\n"
      "XB_SYN_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
XB_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(kSyn2Warps * 32, 1)
qmf_synth_hq_tma_kernel(const __grid_constant__ QmfSynthArgs p, const __grid_constant__ SynTw tw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Syn2Smem &sm = *reinterpret_cast<Syn2Smem *>(smem_raw);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  Syn2Warp &W = sm.w[warp];
  const u32 bar = smem_u32(&sm.bar[warp]);
  if (lane == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();

  const long long warps_total = (long long)gridDim.x * kSyn2Warps;
  i32 *const row = W.rows + (kHist + lane) * kRowW;                 // phase A: this lane's slot
  int4 *const rows4 = reinterpret_cast<int4 *>(W.rows);             // phase B: row -9 first
  const u32 row_s = smem_u32(row);
  const i32 *c32 = reinterpret_cast<const i32 *>(p.rom + kSynV2CoefOffset);

  auto next_unit = [&](long long u) {
    while (u < p.n_units && p.gate && p.gate[u] == 0) u += warps_total;
    return u;
  };
  auto issue = [&](long long u) {  // one 512-byte row per lane
    if (lane == 0) mbar_arrive_expect_tx(bar, 32 * 512);
    bulk_g2s(row_s, p.matrix + u * p.mat_stride + 128 * lane, 512, bar);
  };

  long long u = next_unit((long long)blockIdx.x * kSyn2Warps + warp);
  if (u < p.n_units) issue(u);
  u32 parity = 0;

  while (u < p.n_units) {
    const long long un = next_unit(u + warps_total);
    if (un < p.n_units) bulk_prefetch_l2(p.matrix + un * p.mat_stride + 128 * lane, 512);
    const int16_t *prm = p.params + u * 8;
    const int ov_lb_scale = prm[0], lb_scale = prm[1], hb_scale = prm[2], st_syn = prm[3];
    const int lsb = prm[4], usb = prm[5], split = prm[6];
    const int off0 = p.pos[2 * u], fpos0 = p.pos[2 * u + 1];
    const int Bw0 = off0 >> 7, fp0 = fpos0 >> 6;
    // qmf_dec.c:914-926, :1055
    const int ov_lb_shift = (st_syn - ov_lb_scale) - 8, lb_shift = (st_syn - lb_scale) - 8;
    const int hb_shift = (st_syn - hb_scale) - 8;
    const FoldK fk = fold_consts(-(st_syn - 3) + 1);

    // ---- while the rows are in flight: history, block-shift table, window coefficients ----
    i32 *st32 = reinterpret_cast<i32 *>(p.states + u * 1280);
    {
      i32 wv[20];
#pragma unroll
      for (int t = 0; t < 20; t++) wv[t] = st32[32 * t + lane];
#pragma unroll
      for (int B = 0; B < 10; B++) history_store(rows4, lane, B, Bw0, wv[2 * B], wv[2 * B + 1]);
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const int band = lane + 32 * k;
      W.sh[0][band] = shift_entry(band < lsb ? ov_lb_shift : (band < usb ? hb_shift : 0));
      W.sh[1][band] = shift_entry(band < lsb ? lb_shift : (band < usb ? hb_shift : 0));
    }
    WinCoef wc;
    window_coefs(wc, c32, lane, Bw0, fp0);
    int16_t *pcm = p.pcm + (p.pcm_unit_stride ? u * p.pcm_unit_stride
                                               : ((p.ch_fac == 1) ? u * 2048 : (u / p.ch_fac) * (2048LL * p.ch_fac) + (u % p.ch_fac)));
    __syncwarp();
    mbar_wait(bar, parity);
    parity ^= 1;

    // ---- phase A: lane = slot ----
    {
      // block shift (env_calc.c:1099) of the lane's row in place + magnitude bounds of the shifted values
      i32 mx = 0, mn = 0;
      shift_row(row, W.sh[lane < split ? 0 : 1], mx, mn);
      // Below 2^fast_bits no add of the modulation can saturate and wrapping adds are bit-identical; if any lane of the
      // warp left the bound (never on decoder data) the unit runs with the reference's saturating adds.
      const i32 lim = (i32)(1u << p.fast_bits);
      const bool bad = (mx >= lim) || (mn < -lim);
      if (__any_sync(0xffffffffu, bad))
        slot_modulate<true>(row, tw, fk, 0);
      else
        slot_modulate<false>(row, tw, fk, p.zero);
    }
    __syncwarp();

    // ---- phase B: lane = output pair ----
    if (Bw0 & 1)
      window_unit<1>(rows4, lane, wc, pcm, p.ch_fac);
    else
      window_unit<0>(rows4, lane, wc, pcm, p.ch_fac);

    // ---- filter state (the last 10 blocks) back to the ring, ring / coefficient positions ----
    {
      int B = Bw0 + 8;  // ring block of slot 22: (Bw0 - 22) mod 10
      if (B >= 10) B -= 10;
#pragma unroll
      for (int r = 22; r < 32; r++) {
        i32 w0, w1;
        state_words(rows4, lane, r, w0, w1);
        st32[32 * (2 * B) + lane] = w0;
        st32[32 * (2 * B + 1) + lane] = w1;
        B = B ? B - 1 : 9;
      }
    }
    if (lane == 0) {
      int off = off0, fpos = fpos0;
      if (off0 >= 0 && off0 < 1280 && fpos0 >= 0 && fpos0 < 640 && (fpos0 & 63) == 0) {  // 32 steps, closed form
        off += 1024;
        if (off >= 1280) off -= 1280;
        fpos += 128;
        if (fpos >= 640) fpos -= 640;
      } else {
        for (int s = 0; s < 32; s++) {
          off -= 128;
          if (off < 0) off += 1280;
          fpos += 64;
          if (fpos == 640) fpos = 0;
        }
      }
      p.pos[2 * u] = (int16_t)off;
      p.pos[2 * u + 1] = (int16_t)fpos;
    }
    fence_proxy_async();
    __syncwarp();
    u = un;
    if (u < p.n_units) issue(u);
  }
}

size_t qmf_synth_table_bytes() { return sizeof(SynTables) + 640 * 4; }

// Host-side construction of the table image from the reference-layout QMF ROM blob (leading bytes of
// ia_qmf_dec_tables_struct).  Returns fast_bits (> 0), or -1 if the tables are not the ones the kernels are built for.
int qmf_synth_build_tables(const uint8_t *qrom, uint8_t *out) {
  SynTables *t = reinterpret_cast<SynTables *>(out);
  const int16_t *w32 = reinterpret_cast<const int16_t *>(qrom + kQRomW32);
  const int32_t *dr = reinterpret_cast<const int32_t *>(qrom + kQRomDigRev2_32);
  const int16_t *sc = reinterpret_cast<const int16_t *>(qrom + kQRomSinCosL64);
  const int16_t *al = reinterpret_cast<const int16_t *>(qrom + kQRomAltSinL64);
  const int16_t *c = reinterpret_cast<const int16_t *>(qrom + kQRomQmfC);
  auto hi = [](int16_t v) { return (int32_t)((uint32_t)(uint16_t)v << 16); };
  for (int n = 0; n < 32; n++) t->pre_tw[n] = make_int2(hi(sc[2 * n]), hi(sc[2 * n + 1]));
  for (int n = 0; n < 16; n++) t->alt_tw[n] = make_int2(hi(al[2 * n]), hi(al[2 * n + 1]));
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 3; j++) t->w1[3 * i + j] = make_int2(hi(w32[6 * i + 2 * j]), hi(w32[6 * i + 2 * j + 1]));
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 3; j++)
      t->w2[3 * i + j] = make_int2(hi(w32[48 + 6 * i + 2 * j]), hi(w32[48 + 6 * i + 2 * j + 1]));
  for (int i = 0; i < 32; i++) t->postmap[i] = -1;
  for (int blk = 0; blk < 4; blk++)
    for (int half = 0; half < 2; half++) {
      int cb = (blk >> 1) * 16 + (blk & 1) * 4 + 8 * half;
      int op = ((dr[blk] >> 2) >> 1) + half;
      if (op < 0 || op + 20 >= 32) return -1;
      t->postmap[op] = cb;
      t->postmap[op + 16] = cb | 256;
      t->postmap[op + 4] = cb + 2;
      t->postmap[op + 20] = (cb + 2) | 256;
    }
  for (int i = 0; i < 32; i++)
    if (t->postmap[i] < 0) return -1;
  // the register-resident modulation of the TMA variant has the digit reversal compiled in; the linear-time window of both
  // kernels needs the 640-periodic qmf_c (coefficient by block age, qmf_synth_core.cuh)
  for (int i = 0; i < 32; i++)
    if (t->postmap[i] != (syn::post_src(i) | (i >= 16 ? 256 : 0))) return -1;
  for (int i = 0; i < 640; i++)
    if (c[i] != c[i + 640]) return -1;
  memcpy(out + kSynV2CoefOffset, c, 640 * 4);
  // no-saturation bound of the window-add accumulation (see file header)
  for (int fpos = 0; fpos < 640; fpos += 64)
    for (int k = 0; k < 64; k++) {
      long long s = 0;
      for (int B = 0; B < 10; B++) s += c[fpos + 64 * B + k] < 0 ? -(long long)c[fpos + 64 * B + k] : c[fpos + 64 * B + k];
      if (s * 32768 + 0x4000 >= 0x7fffffffLL) return -1;
    }
  // Largest k such that inputs bounded by 2^k cannot make any add of the modulation saturate.
  // |mul32x16(x, w)| <= |x||w|/65536 + 1.  S_* = max(|sin| + |cos|) over each twiddle table.
  auto smax = [](const int16_t *tab, int pairs) {
    long long m = 0;
    for (int i = 0; i < pairs; i++) {
      long long a = tab[2 * i] < 0 ? -(long long)tab[2 * i] : tab[2 * i];
      long long b = tab[2 * i + 1] < 0 ? -(long long)tab[2 * i + 1] : tab[2 * i + 1];
      if (a + b > m) m = a + b;
    }
    return (double)m;
  };
  const double S_pre = smax(sc, 32), S_w = smax(w32, 30), S_alt = smax(al, 16);
  int fast_bits = 0;
  for (int k = 30; k >= 1; k--) {
    double A = (double)(1ULL << k);
    double B = A * S_pre / 65536.0 + 2.0;                              // pre-twiddle
    for (int stage = 0; stage < 2; stage++) {                          // two radix-4 stages
      double sum4 = 4.0 * B;
      double tw = 2.0 * (sum4 * S_w / 65536.0 + 2.0);
      B = sum4 > tw ? sum4 : tw;
    }
    double F = 2.0 * B;                                                // radix-2
    double G = F * S_alt / 65536.0 + 2.0;                              // post-twiddle
    if (F / 2.0 + 1.0 > G) G = F / 2.0 + 1.0;
    double fold = 2.0 * G;                                             // fold add/sub
    if (fold * 1.0001 + 16.0 < 2147483647.0) {
      fast_bits = k;
      break;
    }
  }
  return fast_bits;
}

void qmf_synth_build_twiddles(const uint8_t *qrom, void *out) {
  SynTw *t = reinterpret_cast<SynTw *>(out);
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
size_t qmf_synth_twiddle_bytes() { return sizeof(SynTw); }

cudaError_t launch_qmf_synth_hq(const QmfSynthArgs &args, int num_sms, cudaStream_t stream) {
  // XAAC_B200_SYNTH_TMA=1 selects the bulk-copy-staged lane = slot variant (bit-identical, 2.09 ms against 1.45 ms per 131 072
  // units: 2 warps per scheduler cannot hide its dependent chains — profiles/r2_synth_tma.md)
  static const bool use_tma = getenv("XAAC_B200_SYNTH_TMA") != nullptr;
  if (!use_tma) {
    static xb::PerDeviceOnce configured_g4;
    const size_t smem4 = sizeof(SynG4Block);
    if (configured_g4.needed()) {
      cudaError_t e = cudaFuncSetAttribute(qmf_synth_hq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
      if (e != cudaSuccess) return e;
      configured_g4.done();
    }
    long long need4 = (args.n_units + kG4Warps - 1) / kG4Warps;
    long long grid4 = num_sms;
    if (grid4 > need4) grid4 = need4;
    if (grid4 < 1) grid4 = 1;
    qmf_synth_hq_kernel<<<(unsigned)grid4, kG4Warps * 32, smem4, stream>>>(args);
    return cudaGetLastError();
  }
  // the rows are staged by 16-byte-granular bulk copies
  if (!args.twiddles || ((uintptr_t)args.matrix & 15) != 0 || (args.mat_stride & 3) != 0) return cudaErrorInvalidValue;
  static xb::PerDeviceOnce configured;
  const size_t smem = sizeof(Syn2Smem);
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(qmf_synth_hq_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.done();
  }
  long long need = (args.n_units + kSyn2Warps - 1) / kSyn2Warps;
  long long grid = num_sms;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  qmf_synth_hq_tma_kernel<<<(unsigned)grid, kSyn2Warps * 32, smem, stream>>>(args, *reinterpret_cast<const SynTw *>(args.twiddles));
  return cudaGetLastError();
}

}  // namespace xb
