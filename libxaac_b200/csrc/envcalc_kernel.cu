// envcalc_kernel.cu — fixed-point complex ("HQ") SBR envelope adjuster for sm_100a (B200).
//
// One warp owns one unit (one frame of one SBR channel); lanes own QMF bands (gain calculation, time-slot adjustment,
// coalesced 128-byte row segments of the matrix) or limiter bands (the sequential (mantissa, exponent) accumulations
// of the noise limiter, whose rounding makes the summation order part of the bit-exact contract).
// Replaces, bit-exactly, the reference stage
//   ixheaacd_calc_sbrenvelope             decoder/ixheaacd_env_calc.c:692-1015            (low_pow_flag = 0)
// and the leaves it calls:
//   ixheaacd_map_sineflags                decoder/ixheaacd_sbrdec_lpfuncs.c:529-560
//   ixheaacd_enery_calc_per_subband_dec   decoder/ixheaacd_env_calc.c:1211-1296  (selector: ixheaacd_enery_calc_per_subband)
//   ixheaacd_enery_calc_persfb            decoder/ixheaacd_env_calc.c:1298-1380
//   ixheaacd_calc_subband_gains / ixheaacd_subbandgain_calc      :616-688, :1382-1452
//   ixheaacd_noiselimiting / ixheaacd_avggain_calc               :229-421, :1454-1562
//   ixheaacd_conv_ergtoamplitude_dec      :450-477                                (selector: ixheaacd_conv_ergtoamplitude)
//   ixheaacd_adapt_noise_gain_calc, ixheaacd_equalize_filt_buff_exp, ixheaacd_filt_buf_update,
//   ixheaacd_noise_level_rescaling        :479-614, :1017-1097
//   ixheaacd_adj_timeslot, ixheaacd_harm_idx_zerotwo / _onethree decoder/ixheaacd_env_dec.c:845-923, env_calc.c:1759-1898
//   ixheaacd_expsubbandsamples_dec, ixheaacd_adjust_scale_dec    :1159-1207, :1099-1157 (selector leaves)
//   ixheaacd_fix_mant_div, ixheaacd_fix_mant_exp_sqrt            decoder/ixheaacd_basic_funcs.c:66-128
//
// Algorithmic HBM bytes per unit (HE-AACv2 tables, 28 generated bands, 32 + 6 slots): high band read + written once
// = 38 x 28 x 8 x 2 = 17 KB, + 1312 B side info + 2 x 464 B state ~= 19.3 KB.
#include <cstdint>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"
#include "sbr_common.cuh"
#include "sbr_glue_units.cuh"
#include "env_common.cuh"

namespace xb {

constexpr int kEnvWarps = 8;
struct EnvWarpS {
  int16_t prm[kEnvPrmWords];
  int16_t st[kEnvStWords];
  int16_t est[2 * kMaxB], gain[2 * kMaxB], noise[2 * kMaxB], sine[2 * kMaxB], orig[2 * kMaxB];
  i32 line[kMaxB];
  int8_t sine_mapped[kMaxB + 8];
};

// POST: the whole-stage driver's variant — after a unit's envelope adjustment the same warp runs sbr_post_unit
// (sbr_dec.c:1205-1245, :1284-1308: previous-frame data, LPC rows, overlap save, synthesis parameters) on the rows it
// has just touched, instead of a separate sbr_post_kernel launch.
template <bool POST>
__device__ __forceinline__ void calc_sbrenvelope_hq_body(const EnvCalcArgs &p, const SbrStageArgs &g) {
  __shared__ EnvRomS rom;
  __shared__ EnvWarpS ws[kEnvWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;
  {
    const int16_t *e = reinterpret_cast<const int16_t *>(p.env_rom);
    const int16_t *m = reinterpret_cast<const int16_t *>(p.misc_rom);
#pragma unroll 1
    for (int i = threadIdx.x; i < 8; i += blockDim.x) rom.lim_gains[i] = e[i];
#pragma unroll 1
    for (int i = threadIdx.x; i < 4; i += blockDim.x) rom.smooth[i] = e[kERomSmooth / 2 + i];
#pragma unroll 1
    for (int i = threadIdx.x; i < 49; i += blockDim.x) rom.inv_int[i] = e[kERomInvInt / 2 + i];
#pragma unroll 1
    for (int i = threadIdx.x; i < 256; i += blockDim.x) rom.inv_table[i] = m[kMRomInvTable / 2 + i];
#pragma unroll 1
    for (int i = threadIdx.x; i < 257; i += blockDim.x) rom.sqrt_table[i] = m[kMRomSqrtTable / 2 + i];
  }
  __syncthreads();
  const i32 *rand_ph = reinterpret_cast<const i32 *>(p.env_rom + kERomRandPh);
  EnvWarpS &w = ws[warp];
  const int warps_total = gridDim.x * kEnvWarps;

  for (long long u = (long long)blockIdx.x * kEnvWarps + warp; u < p.n_units; u += warps_total) {
    __syncwarp();
    if (p.gate && p.gate[u * p.gate_stride] == 0) {
      if (lane == 0 && p.err) p.err[u] = 0;
      if (POST) sbr_post_unit(g, u, lane, false);
      continue;
    }
    {
      const i32 *src = reinterpret_cast<const i32 *>(p.params + u * p.prm_stride);
      // all of a lane's loads of both records in flight before the first store (11 + 4 requests)
      const i32 *ss = reinterpret_cast<const i32 *>(p.state + u * kEnvStWords);
      constexpr int nA = (kEnvPrmWords / 2 + 31) / 32, nB = (kEnvStWords / 2 + 31) / 32;
      i32 va[nA], vb[nB];
#pragma unroll
      for (int q = 0; q < nA; q++) va[q] = lane + 32 * q < kEnvPrmWords / 2 ? __ldg(src + lane + 32 * q) : 0;
#pragma unroll
      for (int q = 0; q < nB; q++) vb[q] = lane + 32 * q < kEnvStWords / 2 ? ss[lane + 32 * q] : 0;
#pragma unroll
      for (int q = 0; q < nA; q++)
        if (lane + 32 * q < kEnvPrmWords / 2) reinterpret_cast<i32 *>(w.prm)[lane + 32 * q] = va[q];
#pragma unroll
      for (int q = 0; q < nB; q++)
        if (lane + 32 * q < kEnvStWords / 2) reinterpret_cast<i32 *>(w.st)[lane + 32 * q] = vb[q];
#pragma unroll 1
      for (int i = lane; i < 2 * kMaxB; i += 32) w.est[i] = w.gain[i] = w.noise[i] = w.sine[i] = w.orig[i] = 0;
#pragma unroll 1
      for (int i = lane; i < kMaxB; i += 32) w.sine_mapped[i] = 8;
    }
    __syncwarp();
    const int16_t *prm = w.prm;
    int16_t *st = w.st;
    i32 *mat = p.matrix + u * (38 * 128);
    int16_t *sf = p.sf + u * 8;
    const int num_env = prm[kEnvNumEnv], trans_env = prm[kEnvTransientEnv];
    const int16_t *border = prm + kEnvBorderVec, *freq_res = prm + kEnvFreqRes, *nborder = prm + kEnvNoiseBorderVec;
    const int num_nf = prm[kEnvNumNfBands];
    const int sb_start = prm[kEnvSubBandStart], sb_end = prm[kEnvSubBandEnd];
    const int max_qmf = prm[kEnvMaxQmfSubband];
    const int max_qmf_prev = p.max_qmf_prev ? p.max_qmf_prev[u * 16] : prm[kEnvMaxQmfSubbandPrev];
    const int num_sub_bands = sb_end - sb_start, skip = max_qmf - sb_start, bands = num_sub_bands - skip;
    const int16_t *fnoise = prm + kEnvFreqNoise;
    const int16_t *sf_arr = prm + kEnvSfArr;
    const int16_t *noise_floor = prm + kEnvNoiseFloor;
    int16_t *filt_me = st + kEnvStFiltMe, *filt_noise = st + kEnvStFiltNoise;
    int16_t *fme = filt_me + 2 * skip, *fno = filt_noise + skip;
    const int sf_hb_in = sf[kSfHb], sf_ov_hb_in = sf[kSfOvHb];

    {  // ixheaacd_map_sineflags: distinct scale-factor bands hit distinct centre bands
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
          w.sine_mapped[q] = old ? 0 : (int8_t)trans_env;
        }
      }
    }
    int adj_e, final_e = 0;
    {  // env_calc.c:772-791
      const int first = (max_qmf_prev > max_qmf ? max_qmf_prev : max_qmf) - sb_start;
      int mx = 0;
#pragma unroll 1
      for (int i = first + lane; i < num_sub_bands; i += 32) mx = max(mx, (int)filt_noise[i]);
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
      const int smooth_len = noise_absc ? 0 : ((1 - prm[kEnvSmoothingMode]) << 2);
      const int input_e = 15 - sf_hb_in;
      const int num_sfb = prm[kEnvNumSfLo + fr];
      const int16_t *ftab = prm + (fr ? kEnvFreqHi : kEnvFreqLo);

      // ---- energy estimation ----
      if (prm[kEnvInterpolFreq]) {
        const i32 inv_width = rom.inv_int[(end - start) >> 0];
#pragma unroll 1
        for (int c = lane; c < sb_end - max_qmf; c += 32) {
          const int k = max_qmf + c;
          i32 max_val = 1;
#pragma unroll 4
          for (int l = start; l < end; l++) {
            max_val = max(max_val, abs_nrm(mat[128 * l + k]));
            max_val = max(max_val, abs_nrm(mat[128 * l + 64 + k]));
          }
          const int pre = pnorm32(max_val) - 4;
          int shift = 16 - pre;
          i32 accu = 0;
#pragma unroll 2
          for (int l = start; l < end; l++) {
            const i32 a = mat[128 * l + k], b = mat[128 * l + 64 + k];
            const i32 ta = sext16(shift > 0 ? (a >> shift) : lsl(a, -shift));
            const i32 tb = sext16(shift > 0 ? (b >> shift) : lsl(b, -shift));
            accu = wadd(accu, wadd(ta * ta, tb * tb));
          }
          if (accu != 0) {
            shift = -pnorm32(accu);
            const i32 sum_m = sext16(shr32_dir_sat_limit(accu, 16 + shift));
            w.est[2 * c] = (int16_t)mult16_shl_sat_(sum_m, inv_width);
            shift -= pre << 1;
            w.est[2 * c + 1] = (int16_t)((input_e << 1) + shift + 1);
          } else {
            w.est[2 * c] = w.est[2 * c + 1] = 0;
          }
        }
      } else {
        // per scale-factor band (env_calc.c:1298-1380); lanes own bands, a band's sfb partners are summed in order
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
#pragma unroll 4
            for (int l = start; l < end; l++) orv |= abs_nrm(mat[128 * l + k]) | abs_nrm(mat[128 * l + 64 + k]);
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
#pragma unroll 2
            for (int l = start; l < end; l++) {
              const i32 ta = sext16(shr32_dir(mat[128 * l + k], s));
              line = add_sat(line, ta * ta);
              const i32 tb = sext16(shr32_dir(mat[128 * l + 64 + k], s));
              line = add_sat(line, tb * tb);
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
              sum_e = ((input_e << 1) + 10) - shift - (pre << 1);
            }
            w.est[2 * (k - first_li)] = (int16_t)sum_m;
            w.est[2 * (k - first_li) + 1] = (int16_t)sum_e;
          }
          __syncwarp();
        }
      }
      if (ftab[0] < sb_start) { err = 1; break; }
      __syncwarp();

      // ---- gains per band (env_calc.c:616-688) ----
      {
        const int f0 = ftab[0], top = ftab[num_sfb];
#pragma unroll 1
        for (int c = lane; c < top - max(max_qmf, f0); c += 32) {
          const int k = max(max_qmf, f0) + c;  // c indexes bands k >= max_qmf in walk order
          int j = 0;
          while (ftab[j + 1] <= k) j++;
          const int li = ftab[j], ui = ftab[j + 1];
          const i32 v = sf_arr[m_off + j];
          const i32 ref_e = sext16((v & 0x3f) - 16), ref_m = sext16(v & 0xffc0);
          bool present = false;
#pragma unroll 1
          for (int kk = li; kk < ui; kk++) present |= (env >= w.sine_mapped[kk - f0]);
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

      // ---- noise limiter: one lane per limiter band (env_calc.c:229-421) ----
      {
        const int16_t *lim = prm + kEnvLimTbl;
        const i32 lg_m = rom.lim_gains[2 * prm[kEnvLimiterGains]], lg_e = rom.lim_gains[2 * prm[kEnvLimiterGains] + 1];
#pragma unroll 1
        for (int c = lane; c < prm[kEnvNumLfBands]; c += 32) {
          const int b0 = lim[c] > skip ? lim[c] - skip : 0, b1 = lim[c + 1] > skip ? lim[c + 1] - skip : 0;
          if (b0 >= b1) continue;
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
#pragma unroll 1
          for (int k = b0; k < b1; k++) {
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
          i32 am = 0, ae = 0;
#pragma unroll 1
          for (int k = b0; k < b1; k++) {
            acc_add(am, ae, ((i32)w.gain[2 * k] * w.est[2 * k]) >> 15, w.gain[2 * k + 1] + w.est[2 * k + 1]);
            if (w.sine[2 * k] != 0) acc_add(am, ae, w.sine[2 * k], w.sine[2 * k + 1]);
            else if (!noise_absc) acc_add(am, ae, w.noise[2 * k], w.noise[2 * k + 1]);
          }
          nv = 16 - norm32(am);
          if (nv > 0) { am >>= nv; ae += nv; }
          i32 bg_m;
          i32 bg_e = sext16(mant_div(sum_m, sext16(am), bg_m, rom));
          bg_e = sext16(bg_e + (sum_e - sext16(ae)) + 1);
          if (bg_e > 2 || (bg_e == 2 && bg_m > 0x5061)) { bg_m = 0x5061; bg_e = 2; }
#pragma unroll 1
          for (int k = b0; k < b1; k++) {
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

      // ---- energies -> amplitudes, start-up / exponent equalisation (env_calc.c:450-477, 495-516, 1017-1058) ----
      int noise_e = sext16(start < 32 ? adj_e : final_e);
      const bool start_up = st[kEnvStStartUp] != 0;
      __syncwarp();
#pragma unroll 1
      for (int k = lane; k < bands; k += 32) {
        mant_exp_sqrt(&w.sine[2 * k], rom);
        mant_exp_sqrt(&w.gain[2 * k], rom);
        mant_exp_sqrt(&w.noise[2 * k], rom);
        const int shift = (noise_e - w.noise[2 * k + 1]) - 4;
        if (shift > 0) w.noise[2 * k] = (int16_t)(w.noise[2 * k] >> min(shift, 31));
        else w.noise[2 * k] = (int16_t)lsl((i32)w.noise[2 * k], min(-shift, 31));
        if (start_up) {
          fme[2 * k] = w.gain[2 * k];
          fme[2 * k + 1] = w.gain[2 * k + 1];
          fno[k] = w.noise[2 * k];
        } else {
          const i32 fe = fme[2 * k + 1], fm = fme[2 * k], diff = w.gain[2 * k + 1] - fe;
          if (diff >= 0) {
            fme[2 * k + 1] = w.gain[2 * k + 1];
            fme[2 * k] = (int16_t)(fm >> (diff & 31));  // x86 shift-count semantics of the pinned reference build
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

      // ---- time-slot adjustment (env_calc.c:518-609, env_dec.c:845-923) ----
      int ph_index = st[kEnvStPhIndex], harm = st[kEnvStHarmIndex];
      int filt_noise_e = st[kEnvStNoiseE];
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
          }
        }
        {
          const int diff = filt_noise_e - noise_e;
          if (diff > 0) for (int k = lane; k < num_sub_bands; k += 32) filt_noise[k] = (int16_t)(filt_noise[k] >> (diff & 31));
          else if (diff < 0) for (int k = lane; k < num_sub_bands; k += 32) filt_noise[k] = (int16_t)lsl((i32)filt_noise[k], (-diff) & 31);
          filt_noise_e = noise_e;
        }
        __syncwarp();
        const i32 ratio = (l - start) < smooth_len ? rom.smooth[l - start] : 0;
        const i32 direct = sat16(0x7fff - ratio);
        const int sc = sext16(sext16(scale_change) - 1);
        const int nfe = sext16(noise_e - 16);
#pragma unroll 1
        for (int k = lane; k < bands; k += 32) {
          i32 g = w.gain[2 * k], nz = w.noise[2 * k];
          if (ratio) {
            const i32 t = sext16(mult16_(ratio, fme[2 * k]) + mult16_(direct, g));
            const i32 t1 = sext16(mult16_(ratio, fno[k]) + mult16_(direct, nz));
            g = sext16(lsl(t, 1));
            nz = sext16(lsl(t1, 1));
            fme[2 * k] = (int16_t)g;
            fno[k] = (int16_t)nz;
          }
          i32 *pr = mat + 128 * l + max_qmf + k, *pi = pr + 64;
          i32 sr = mul32x16(*pr, g), si = mul32x16(*pi, g);
          const int shift = sext16(w.gain[2 * k + 1] - sc);
          if (shift > 0) { sr = shl32(sr, shift); si = shl32(si, shift); }
          else { sr = shr32(sr, -shift); si = shr32(si, -shift); }
          const i32 sm = w.sine[2 * k];
          if (sm != 0) {
            const int t = sext16(w.sine[2 * k + 1] - nfe);
            i32 sl;
            if (t > 0) sl = shl32(sm, t);
            else if (harm & 1) sl = shr32(sm, -t);
            else sl = shr32(sm, t);  // env_calc.c:1796 passes the non-positive count unnegated
            if (harm == 0) sr = add_sat(sr, sl);
            else if (harm == 2) sr = sub_sat(sr, sl);
            else {
              const bool finv = (((max_qmf + k) & 1) != 0) != (harm == 1);
              si = finv ? add_sat(si, sl) : sub_sat(si, sl);
            }
          } else if (!noise_absc) {
            const i32 rv = __ldg(rand_ph + ph_index + k + 1);
            const i32 prd = (rv >> 16) * nz, pid = sext16(rv) * nz;
            sr = add_sat(sr, prd == 0x40000000 ? 0x7fffffff : shl32(prd, 1));
            si = add_sat(si, pid == 0x40000000 ? 0x7fffffff : shl32(pid, 1));
          }
          *pr = sr;
          *pi = si;
        }
        ph_index = (ph_index + bands) & 511;
        harm = (harm + 1) & 3;
        __syncwarp();
      }
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

    if (!err) {  // env_calc.c:956-1013
      const int first_start = border[0] * 2;
      int ov_reserve = 0, reserve = 0;
      __syncwarp();
      if (prm[kEnvChannelMode] == 3) {
        ov_reserve = warp_headroom(mat, max_qmf, sb_end, 0, first_start, lane);
        reserve = warp_headroom(mat, max_qmf, sb_end, first_start, 32, lane);
      }
      const int ov_adj_e = 15 - sf_ov_hb_in;
      const int output_e = max(ov_adj_e - ov_reserve, adj_e - reserve);
      warp_adjust_scale(mat, max_qmf, sb_end, 0, first_start, ov_adj_e - output_e, lane);
      warp_adjust_scale(mat, max_qmf, sb_end, first_start, prm[kEnvNumTimeSlots] * prm[kEnvTimeStep], adj_e - output_e,
                        lane);
      if (lane == 0) {
        sf[kSfHb] = (int16_t)(15 - output_e);
        sf[kSfOvHb] = (int16_t)(15 - final_e);
        st[kEnvStTransPrev] = (trans_env == num_env) ? 0 : -1;
      }
    }
    __syncwarp();
    {
      i32 *ds = reinterpret_cast<i32 *>(p.state + u * kEnvStWords);
#pragma unroll 1
      for (int i = lane; i < kEnvStWords / 2; i += 32) ds[i] = reinterpret_cast<const i32 *>(w.st)[i];
      if (lane == 0 && p.err) p.err[u] = err ? (i32)0x80000000 : 0;
    }
    if (POST) {
      __syncwarp();
      sbr_post_unit(g, u, lane, __shfl_sync(full, err, 0) != 0);
    }
  }
}

__global__ void __launch_bounds__(kEnvWarps * 32)
calc_sbrenvelope_hq_kernel(EnvCalcArgs p) {
  calc_sbrenvelope_hq_body<false>(p, SbrStageArgs());
}

__global__ void __launch_bounds__(kEnvWarps * 32)
calc_sbrenvelope_hq_post_kernel(EnvCalcArgs p, SbrStageArgs g) {
  calc_sbrenvelope_hq_body<true>(p, g);
}

cudaError_t launch_calc_sbrenvelope_hq(const EnvCalcArgs &args, int num_sms, cudaStream_t stream) {
  long long need = (args.n_units + kEnvWarps - 1) / kEnvWarps;
  long long grid = (long long)num_sms * 4;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  calc_sbrenvelope_hq_kernel<<<(unsigned)grid, kEnvWarps * 32, 0, stream>>>(args);
  return cudaGetLastError();
}

cudaError_t launch_calc_sbrenvelope_hq_post(const EnvCalcArgs &args, const SbrStageArgs &g, int num_sms, cudaStream_t stream) {
  long long need = (args.n_units + kEnvWarps - 1) / kEnvWarps;
  long long grid = (long long)num_sms * 4;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  calc_sbrenvelope_hq_post_kernel<<<(unsigned)grid, kEnvWarps * 32, 0, stream>>>(args, g);
  return cudaGetLastError();
}

}  // namespace xb
