// sbr_sideinfo_kernel.cu — SBR side-info dequantisation for sm_100a (B200): SURVEY.md §8f-3.
//
// One warp owns one element (one or two SBR channels of one frame); lanes own scale-factor bands.  Replaces, bit-exactly,
// on the fixed-point path (usac_flag = enh_sbr = 0, ec_flag = 0, ldmps_present = 0, not ELD):
//   ixheaacd_dec_sbrdata                     decoder/ixheaacd_env_dec.c:628-725
//   ixheaacd_dec_envelope                    :727-843   (timing check, concealment, one retry after a failed range check)
//   ixheaacd_lean_sbrconcealment             decoder/ixheaacd_sbrdec_lpfuncs.c:196-253
//   ixheaacd_wrong_timing_compensate         env_dec.c:238-285
//   ixheaacd_process_del_cod_env_data        :122-236   + ixheaacd_map_res_energy :88-120
//   ixheaacd_check_env_data                  :287-320
//   ixheaacd_dequant_env_data                :322-342
//   ixheaacd_calc_noise_floor                :396-494   + ixheaacd_limit_noise_floor_fac :344-376
//   ixheaacd_sbr_env_dequant_coup_fix        :516-584   + ixheaacd_fix_mant_exp_add / ixheaacd_fix_mant_div (basic_funcs.c:35, :66)
// The reference walks every envelope band by band; here the frequency-direction delta decoding is a warp prefix sum (the
// WORD16 adds wrap, so the sum is associative), the time-direction decoding and the low -> high resolution mapping of the
// previous-frame energies are independent per band.  Envelopes stay sequential (each one updates sfb_nrg_prev for the next).
// The element record (2608 bytes, include/xaac_b200.h XAAC_SD_*) lives in shared memory while the warp works on it.
// Algorithmic HBM bytes per element: 2 x 2608.
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/xaac_b200.h"
#include "fixmath.cuh"
#include "kernels.h"

namespace xb {

constexpr int kSdWarps = 8;
constexpr int kMRomLogDual = 1042;  // WORD16 log_dual_is_table[65] inside ixheaacd_misc_tables

struct SdCtx {
  int16_t *hdr;             // channel block holding this channel's err_flag / err_flag_prev (channel 0's when the header is shared)
  int16_t *ch;              // this channel's block
  int16_t *other;           // the other channel's block (its sfb_nrg_prev is read when the coupling mode changed)
  int16_t *save;            // 56 words of scratch
  const int16_t *log_dual;  // misc tables
  int lane;
};

XB_DEV i32 s16(i32 v) { return (i32)(int16_t)v; }

// ixheaacd_map_res_energy: energy of band `index` of an envelope of resolution `res` into the high-resolution previous-frame row
XB_DEV void map_res_energy(i32 val, int16_t *prev, int drc, int index, int res) {
  if (res == 0) {
    if (drc >= 0) {
      if (index < drc) {
        prev[index] = (int16_t)val;
      } else {
        const int i2 = 2 * index - drc;
        prev[i2] = (int16_t)val;
        prev[i2 + 1] = (int16_t)val;
      }
    } else {
      const int o = -drc;
      if (index < o) {
        prev[3 * index] = (int16_t)val;
        prev[3 * index + 1] = (int16_t)val;
        prev[3 * index + 2] = (int16_t)val;
      } else {
        const int i2 = 2 * index + o;
        prev[i2] = (int16_t)val;
        prev[i2 + 1] = (int16_t)val;
      }
    }
  } else {
    prev[index] = (int16_t)val;
  }
}

// ixheaacd_process_del_cod_env_data: lanes own bands lane and lane + 32 of the envelope being decoded
XB_DEV void del_cod_env(const SdCtx &c) {
  const unsigned full = 0xffffffffu;
  int16_t *v = c.ch + XAAC_SDC_ENV, *prev = c.ch + XAAC_SDC_PREV_NRG;
  const int nlo = c.ch[XAAC_SDC_NUM_SF_LO], nhi = c.ch[XAAC_SDC_NUM_SF_HI];
  const int num_env = c.ch[XAAC_SDC_NUM_ENV];
  int drc = 2 * nlo - nhi;  // the reference flips its sign for good inside the first low-resolution time-direction envelope
  int off = 0;
  for (int i = 0; i < num_env; i++) {
    const int dir = c.ch[XAAC_SDC_DIR + (i & 7)], res = c.ch[XAAC_SDC_FREQ_RES + (i & 7)];
    const int nsb = res ? nhi : nlo;
    const int b0 = c.lane, b1 = c.lane + 32;
    const bool in0 = b0 < nsb && off + b0 < 448, in1 = b1 < nsb && off + b1 < 448;
    if (dir == 0) {  // frequency direction: running sum over the bands, every partial sum also refreshes sfb_nrg_prev
      i32 x0 = in0 ? v[off + b0] : 0, x1 = in1 ? v[off + b1] : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const i32 y0 = __shfl_up_sync(full, x0, d), y1 = __shfl_up_sync(full, x1, d);
        if (c.lane >= d) { x0 += y0; x1 += y1; }
      }
      x1 += __shfl_sync(full, x0, 31);
      x0 = s16(x0);
      x1 = s16(x1);
      if (in0) { v[off + b0] = (int16_t)x0; map_res_energy(x0, prev, drc, b0, res); }
      if (in1) { v[off + b1] = (int16_t)x1; map_res_energy(x1, prev, drc, b1, res); }
    } else if (res == 0) {  // time direction, low resolution: the previous energies sit in the high-resolution row
      if (drc < 0) {
        drc = -drc;
        const int tar = min(drc, nsb);
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int b = h ? b1 : b0;
          if (!(h ? in1 : in0)) continue;
          if (b < tar) {
            const i32 t = s16(v[off + b] + prev[3 * b]);
            prev[3 * b] = prev[3 * b + 1] = prev[3 * b + 2] = (int16_t)t;
            v[off + b] = (int16_t)t;
          } else {
            const int i3 = 2 * b + drc;
            const i32 t = s16(v[off + b] + prev[i3]);
            prev[i3] = prev[i3 + 1] = (int16_t)t;
            v[off + b] = (int16_t)t;
          }
        }
      } else {
        const int tar = min(drc, nsb);
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int b = h ? b1 : b0;
          if (!(h ? in1 : in0)) continue;
          if (b < tar) {
            const i32 t = s16(v[off + b] + prev[b]);
            v[off + b] = (int16_t)t;
            prev[b] = (int16_t)t;
          } else {
            const int i2 = b < drc ? b : 2 * b - drc;
            const i32 t = s16(v[off + b] + prev[i2]);
            prev[i2] = prev[i2 + 1] = (int16_t)t;
            v[off + b] = (int16_t)t;
          }
        }
      }
    } else {  // time direction, high resolution
      if (in0) { const i32 t = s16(v[off + b0] + prev[b0]); v[off + b0] = (int16_t)t; prev[b0] = (int16_t)t; }
      if (in1) { const i32 t = s16(v[off + b1] + prev[b1]); v[off + b1] = (int16_t)t; prev[b1] = (int16_t)t; }
    }
    off += nsb;
    __syncwarp();
  }
}

// ixheaacd_lean_sbrconcealment
XB_DEV void conceal(const SdCtx &c) {
  int16_t *ch = c.ch;
  const int nts = ch[XAAC_SDC_NUM_TIME_SLOTS];
  const int nhi = ch[XAAC_SDC_NUM_SF_HI];
  const int coupling = ch[XAAC_SDC_PREV_COUPLING];
  i32 target = coupling == 2 ? 12 : 0, step = 1;
  if (ch[XAAC_SDC_HDR_AMP_RES] == 0) { target <<= 1; step <<= 1; }
  __syncwarp();
  for (int i = c.lane; i < nhi && i < 448; i += 32)
    ch[XAAC_SDC_ENV + i] = (int16_t)(ch[XAAC_SDC_PREV_NRG + (i < 56 ? i : 55)] > target ? -step : step);
  for (int i = c.lane; i < 56; i += 32) ch[XAAC_SDC_ADD_HARM + i] = 0;
  if (c.lane < 5) ch[XAAC_SDC_INVF + c.lane] = ch[XAAC_SDC_PREV_INVF + c.lane];  // sizeof(WORD32) * MAX_INVF_BANDS
  if (c.lane < 10) ch[XAAC_SDC_NOISE + c.lane] = 0;
  if (c.lane == 0) {
    ch[XAAC_SDC_AMP_RES] = ch[XAAC_SDC_PREV_AMP_RES];
    ch[XAAC_SDC_COUPLING] = (int16_t)coupling;
    ch[XAAC_SDC_MAX_QMF_SB] = ch[XAAC_SDC_PREV_MAX_QMF];
    ch[XAAC_SDC_NUM_ENV] = 1;
    const int16_t start = (int16_t)(ch[XAAC_SDC_PREV_END_POS] - nts);
    ch[XAAC_SDC_BORDER] = start;
    ch[XAAC_SDC_BORDER + 1] = (int16_t)nts;
    ch[XAAC_SDC_NOISE_BORDER] = start;
    ch[XAAC_SDC_NOISE_BORDER + 1] = (int16_t)nts;
    ch[XAAC_SDC_FREQ_RES] = 1;
    ch[XAAC_SDC_TRANSIENT_ENV] = -1;
    ch[XAAC_SDC_NUM_NOISE_ENV] = 1;
    ch[XAAC_SDC_NUM_ENV_SFAC] = (int16_t)nhi;
    ch[XAAC_SDC_DIR] = 1;
    ch[XAAC_SDC_DIR_NOISE] = 1;
  }
  __syncwarp();
}

// ixheaacd_wrong_timing_compensate; false: the estimated start position is negative (the reference returns -1)
XB_DEV bool timing_compensate(const SdCtx &c) {
  int16_t *ch = c.ch;
  i32 start = ch[XAAC_SDC_PREV_END_POS] - ch[XAAC_SDC_NUM_TIME_SLOTS];
  const i32 b0 = ch[XAAC_SDC_BORDER], b1 = ch[XAAC_SDC_BORDER + 1];
  const i32 ref_len = b1 - b0;
  i32 new_len = b1 - start;
  if (new_len <= 0) {
    new_len = ref_len;
    start = b0;
  }
  i32 delta = s16(c.log_dual[min(max(ref_len, 0), 64)] - c.log_dual[min(max(new_len, 0), 64)]);
  delta = s16(delta >> (13 - ch[XAAC_SDC_AMP_RES]));
  __syncwarp();
  if (c.lane == 0) {
    ch[XAAC_SDC_BORDER] = (int16_t)start;
    ch[XAAC_SDC_NOISE_BORDER] = (int16_t)start;
  }
  __syncwarp();
  if (start < 0) return false;
  if (ch[XAAC_SDC_COUPLING] != 2) {
    const int n = ch[XAAC_SDC_FREQ_RES] ? ch[XAAC_SDC_NUM_SF_HI] : ch[XAAC_SDC_NUM_SF_LO];
    for (int i = c.lane; i < n && i < 448; i += 32) ch[XAAC_SDC_ENV + i] = (int16_t)(ch[XAAC_SDC_ENV + i] + delta);
  }
  __syncwarp();
  return true;
}

// ixheaacd_check_env_data; true: a value above the amplitude-resolution maximum was found
XB_DEV bool check_env(const SdCtx &c) {
  int16_t *ch = c.ch;
  const i32 mx = 70 >> ch[XAAC_SDC_AMP_RES];
  const int n = min((int)ch[XAAC_SDC_NUM_ENV_SFAC], 448), nhi = min((int)ch[XAAC_SDC_NUM_SF_HI], 56);
  bool bad = false;
  for (int i = c.lane; i < n; i += 32) {
    const i32 x = ch[XAAC_SDC_ENV + i];
    bad |= x > mx;
    if (x < 0) ch[XAAC_SDC_ENV + i] = 0;
  }
  for (int i = c.lane; i < nhi; i += 32) {
    const i32 x = ch[XAAC_SDC_PREV_NRG + i];
    ch[XAAC_SDC_PREV_NRG + i] = (int16_t)(x < 0 ? 0 : (x > mx ? mx : x));
  }
  __syncwarp();
  return __any_sync(0xffffffffu, bad);
}

// ixheaacd_dequant_env_data
XB_DEV void dequant_env(const SdCtx &c) {
  int16_t *ch = c.ch;
  const int a1 = 1 - ch[XAAC_SDC_AMP_RES];
  const int n = min((int)ch[XAAC_SDC_NUM_ENV_SFAC], 448);
  for (int i = c.lane; i < n; i += 32) {
    i32 e = ch[XAAC_SDC_ENV + i];
    const i32 mant = (e & a1) ? 0x5a80 : 0x4000;
    e = (e >> a1) + 23;
    ch[XAAC_SDC_ENV + i] = (int16_t)(mant | (e & 63));
  }
  __syncwarp();
}

// ixheaacd_dec_envelope; returns 0, 1 (IA_FATAL_ERROR) or 2 (the -1 of the timing compensation)
XB_DEV i32 dec_envelope(const SdCtx &c) {
  int16_t *ch = c.ch, *hdr = c.hdr;
  for (int attempt = 0; attempt < 2; attempt++) {  // the reference calls itself once more after a failed range check
    i32 t1 = ch[XAAC_SDC_PREV_END_POS] - ch[XAAC_SDC_NUM_TIME_SLOTS];
    if (t1 < 0) return 1;  // IA_FATAL_ERROR
    t1 = ch[XAAC_SDC_BORDER] - t1;
    int err = hdr[XAAC_SDC_ERR_FLAG], err_prev = hdr[XAAC_SDC_ERR_FLAG_PREV];
    if (!err_prev && !err && t1 != 0) {
      if (ch[XAAC_SDC_DIR] == 1) err = 1;
      else err_prev = 1;
      __syncwarp();
      if (c.lane == 0) {
        hdr[XAAC_SDC_ERR_FLAG] = (int16_t)err;
        hdr[XAAC_SDC_ERR_FLAG_PREV] = (int16_t)err_prev;
      }
      __syncwarp();
    }
    if (err) {
      conceal(c);
      del_cod_env(c);
      break;
    }
    const int num = min((int)ch[XAAC_SDC_NUM_SF_HI], 56);
    if (err_prev) {
      if (!timing_compensate(c)) return 2;  // the reference's -1
      const int cm = ch[XAAC_SDC_COUPLING], pcm = ch[XAAC_SDC_PREV_COUPLING];
      if (cm != pcm) {
        int16_t *p0 = ch + XAAC_SDC_PREV_NRG, *p1 = c.other + XAAC_SDC_PREV_NRG;
        if (pcm == 2) {
          for (int i = c.lane; i < num; i += 32) p0[i] = p1[i];
        } else if (cm == 1) {
          for (int i = c.lane; i < num; i += 32) p0[i] = (int16_t)(((i32)p0[i] + (i32)p1[i]) >> 1);
        } else if (cm == 2) {
          // memset(.., SBR_ENERGY_PAN_OFFSET, sizeof(WORD16) * num): every BYTE becomes 12
          for (int i = c.lane; i < num; i += 32) p0[i] = (int16_t)0x0c0c;
        }
        __syncwarp();
      }
    }
    for (int i = c.lane; i < 56; i += 32) c.save[i] = ch[XAAC_SDC_PREV_NRG + i];
    __syncwarp();
    del_cod_env(c);
    if (!check_env(c)) break;
    if (c.lane == 0) hdr[XAAC_SDC_ERR_FLAG] = 1;
    for (int i = c.lane; i < 56; i += 32) ch[XAAC_SDC_PREV_NRG + i] = c.save[i];
    __syncwarp();
  }
  dequant_env(c);
  return 0;
}

// ixheaacd_calc_noise_floor (at most 10 values: one lane)
XB_DEV i32 calc_noise_floor(const SdCtx &c) {
  int16_t *ch = c.ch;
  i32 rc = 0;
  __syncwarp();
  if (c.lane == 0) {
    int16_t *nf = ch + XAAC_SDC_NOISE, *pn = ch + XAAC_SDC_PREV_NOISE;
    const int nnf = min(max((int)ch[XAAC_SDC_NUM_NF], 0), 5), nne = ch[XAAC_SDC_NUM_NOISE_ENV];
    if (ch[XAAC_SDC_DIR_NOISE] == 0) {
      for (int i = 1; i < nnf; i++) nf[i] = (int16_t)(nf[i] + nf[i - 1]);
    } else {
      for (int i = 0; i < nnf; i++) nf[i] = (int16_t)(nf[i] + pn[i]);
    }
    if (nne > 1) {
      if (ch[XAAC_SDC_DIR_NOISE + 1] == 0) {
        for (int i = 1; i < nnf; i++) nf[nnf + i] = (int16_t)(nf[nnf + i] + nf[nnf + i - 1]);
      } else {
        for (int i = 0; i < nnf; i++) nf[nnf + i] = (int16_t)(nf[nnf + i] + nf[i]);
      }
    }
    const int tot = min(max(nne * nnf, 0), 10);
    for (int i = 0; i < tot; i++) nf[i] = (int16_t)min(max((i32)nf[i], 0), 35);
    const int o = nnf * (nne - 1);
    if (o < 0 || o >= 10) {
      rc = 1;  // IA_FATAL_ERROR
    } else {
      for (int i = 0; i < nnf; i++) pn[i] = nf[o + i];
      if (ch[XAAC_SDC_COUPLING] != 2)
        for (int i = 0; i < tot; i++) nf[i] = (int16_t)(0x4000 + ((6 + 1 + 38 - nf[i]) & 63));
    }
  }
  rc = __shfl_sync(0xffffffffu, rc, 0);
  __syncwarp();
  return rc;
}

// ixheaacd_fix_mant_exp_add / ixheaacd_fix_mant_div on WORD16 (mantissa, exponent) pairs
XB_DEV void mant_exp_add(i32 m1, i32 e1, i32 m2, i32 e2, i32 &rm, i32 &re) {
  i32 ne = e1 - e2;
  if (ne < 0) {
    m1 = s16(m1) >> min(-ne, 31);
    ne = e2;
  } else {
    m2 = s16(m2) >> min(ne, 31);
    ne = e1;
  }
  i32 nm = m1 + m2;
  if ((nm < 0 ? -nm : nm) >= 0x8000) {
    nm >>= 1;
    ne++;
  }
  rm = s16(nm);
  re = s16(ne);
}
XB_DEV i32 mant_div16(i32 op1, i32 op2, i32 &rm, const int16_t *inv_table) {
  const int pre = norm32(op2) - 16;
  int index = (lsl(op2, pre) >> 5) & 511;
  int post;
  if (index == 0) {
    post = norm32(op1) - 16;
    rm = s16(lsl(op1, post));
  } else {
    const i32 ratio = (i32)inv_table[(index - 1) >> 1] * op1;
    post = norm32(ratio) - 1;
    rm = s16(lsl(ratio, post) >> 15);
  }
  return pre - post;
}

// ixheaacd_sbr_env_dequant_coup_fix: level / balance pair -> left / right energies
XB_DEV void coup_fix(int16_t *l, int16_t *r, int lane, const int16_t *inv_table) {
  const int n = min((int)l[XAAC_SDC_NUM_ENV_SFAC], 448);
  for (int i = lane; i < n; i += 32) {
    const i32 rv = r[XAAC_SDC_ENV + i], lv = l[XAAC_SDC_ENV + i];
    const i32 rmant = s16(rv & 0xffc0), rexp = s16((rv & 63) - 34);
    const i32 lmant = s16(lv & 0xffc0), lexp = s16((lv & 63) - 16);
    i32 pm, pe, nrm;
    mant_exp_add(rmant, rexp, 0x4000, 1, pm, pe);
    i32 nre = s16(mant_div16(lmant, pm, nrm, inv_table));
    nre = s16(nre + lexp - pe + 2);
    const i32 nlm = s16((rmant * nrm) >> 15);  // ixheaac_mult16_shl
    const i32 nle = s16(rexp + nre);
    r[XAAC_SDC_ENV + i] = (int16_t)(((nrm + 32) & 0xffc0) + ((nre + 16) & 63));
    l[XAAC_SDC_ENV + i] = (int16_t)(((nlm + 32) & 0xffc0) + ((nle + 16) & 63));
  }
  const int i_end = min(max(l[XAAC_SDC_NUM_NF] * l[XAAC_SDC_NUM_NOISE_ENV], 0), 10);
  if (lane < i_end) {
    const i32 lexp = s16((l[XAAC_SDC_NOISE + lane] & 63) - 38), rexp = s16(r[XAAC_SDC_NOISE + lane] - 12);
    i32 pm, pe, nrm;
    mant_exp_add(0x4000, s16(1 + rexp), 0x4000, 1, pm, pe);
    i32 nre = s16(mant_div16(0x4000, pm, nrm, inv_table));
    nre = s16(nre + lexp - pe + 2);
    const i32 nle = s16(nre + rexp);
    r[XAAC_SDC_NOISE + lane] = (int16_t)(((nrm + 32) & 0xffc0) + ((nre + 38) & 63));
    l[XAAC_SDC_NOISE + lane] = (int16_t)(((nrm + 32) & 0xffc0) + ((nle + 38) & 63));
  }
  __syncwarp();
}

struct SdWarpS {
  int16_t rec[XAAC_SD_WORDS];
  int16_t save[64];
};

__global__ void __launch_bounds__(kSdWarps * 32) sbr_sideinfo_kernel(int16_t *records, long long n, const uint8_t *misc_rom) {
  __shared__ __align__(16) SdWarpS ws[kSdWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  SdWarpS &w = ws[warp];
  const int16_t *log_dual = reinterpret_cast<const int16_t *>(misc_rom + kMRomLogDual);
  const int16_t *inv_table = reinterpret_cast<const int16_t *>(misc_rom + kMRomInvTable);
  const long long warps_total = (long long)gridDim.x * kSdWarps;
  for (long long u = (long long)blockIdx.x * kSdWarps + warp; u < n; u += warps_total) {
    int4 *g = reinterpret_cast<int4 *>(records + u * XAAC_SD_WORDS);
    int4 *s = reinterpret_cast<int4 *>(w.rec);
    __syncwarp();
    for (int i = lane; i < XAAC_SD_WORDS / 8; i += 32) s[i] = g[i];
    __syncwarp();
    int16_t *c0 = w.rec + XAAC_SD_CH, *c1 = c0 + XAAC_SD_CH_WORDS;
    const bool two = w.rec[XAAC_SD_NUM_CH] == 2, shared = w.rec[XAAC_SD_SHARED_HDR] != 0;
    SdCtx a{c0, c0, c1, w.save, log_dual, lane};
    SdCtx b{shared ? c0 : c1, c1, c0, w.save, log_dual, lane};
    i32 rc = dec_envelope(a);
    if (rc == 0) rc = calc_noise_floor(a);
    if (rc == 0 && two) {
      const int error_code = c0[XAAC_SDC_ERR_FLAG];
      rc = dec_envelope(b);
      if (rc == 0) rc = calc_noise_floor(b);
      if (rc == 0 && !error_code && c0[XAAC_SDC_ERR_FLAG]) rc = dec_envelope(a);
      if (rc == 0 && c0[XAAC_SDC_COUPLING]) {
        __syncwarp();
        if (lane == 0) c0[XAAC_SDC_NUM_NOISE_SFAC] = (int16_t)(c1[XAAC_SDC_NUM_NF] * c1[XAAC_SDC_NUM_NOISE_ENV]);
        coup_fix(c0, c1, lane, inv_table);
      }
    }
    __syncwarp();
    if (lane == 0) w.rec[XAAC_SD_ERR] = (int16_t)rc;
    __syncwarp();
    for (int i = lane; i < XAAC_SD_WORDS / 8; i += 32) g[i] = s[i];
  }
}

// ---- PS side info: ixheaacd_decode_ps_data (decoder/ixheaacd_ps_bitdec.c:98-282) -------------------------------------------------
// One warp per PS instance, the record (1152 bytes) in shared memory.  The delta chains are clamped after every step, so they are
// sequential by contract; the IID and ICC chains are independent of each other and run side by side on two lanes, the copies and
// the 34 -> 20 band mapping on all lanes.
XB_DEV i32 ps_clamp(i32 v, i32 lo, i32 hi) { return v < lo ? lo : (v > hi ? hi : v); }
XB_DEV i32 ps_div2(i32 op) { return s16(op < 0 ? -((-op) >> 1) : (op >> 1)); }
XB_DEV i32 ps_div3(i32 op) {  // ixheaacd_divideby3 (ps_bitdec.c:77)
  const bool sign = op < 0;
  if (sign) op = -op;
  const i32 t = s16((s16(op << 2) * 0x2aab) >> 15);
  const i32 r = t >> 2;
  return s16(sign ? -r : r);
}
// one parameter set (IID or ICC) of one envelope: delta decoding + clamps + 10 -> 20 expansion (ps_bitdec.c:126-196)
XB_DEV void ps_decode_set(int16_t *tab, const int16_t *prev, bool enable, bool dt, int mode, i32 lo, i32 hi) {
  const int nb = mode == 0 ? 10 : (mode == 1 ? 20 : 34), step = mode ? 1 : 2;
  if (enable) {
    if (dt) {
      for (int i = 0; i < nb; i++) tab[i] = (int16_t)ps_clamp(s16(prev[step * i] + tab[i]), lo, hi);
    } else {
      tab[0] = (int16_t)ps_clamp(tab[0], lo, hi);
      for (int i = 1; i < nb; i++) tab[i] = (int16_t)ps_clamp(s16(tab[i - 1] + tab[i]), lo, hi);
    }
  } else {
    for (int i = 0; i < nb; i++) tab[i] = 0;
  }
  if (step == 2)
    for (int i = 2 * nb - 1; i != 0; i--) tab[i] = tab[i >> 1];
}
// ixheaacd_map_34_params_to_20 (sbrdec_lpfuncs.c:561): in place, every output reads inputs at or above its own index
XB_DEV void ps_map_34_to_20(int16_t *p) {
  p[0] = (int16_t)ps_div3(p[0] + p[0] + p[1]);
  p[1] = (int16_t)ps_div3(p[1] + p[2] + p[2]);
  p[2] = (int16_t)ps_div3(p[3] + p[3] + p[4]);
  p[3] = (int16_t)ps_div3(p[4] + p[5] + p[5]);
  p[4] = (int16_t)ps_div2(p[6] + p[7]);
  p[5] = (int16_t)ps_div2(p[8] + p[9]);
  p[6] = p[10];
  p[7] = p[11];
  p[8] = (int16_t)ps_div2(p[12] + p[13]);
  p[9] = (int16_t)ps_div2(p[14] + p[15]);
  p[10] = p[16];
  p[11] = p[17];
  p[12] = p[18];
  p[13] = p[19];
  p[14] = (int16_t)ps_div2(p[20] + p[21]);
  p[15] = (int16_t)ps_div2(p[22] + p[23]);
  p[16] = (int16_t)ps_div2(p[24] + p[25]);
  p[17] = (int16_t)ps_div2(p[26] + p[27]);
  p[18] = (int16_t)ps_div2(ps_div2(p[28] + p[29] + p[30] + p[31]));
  p[19] = (int16_t)ps_div2(p[32] + p[33]);
}

constexpr int kPsdWarps = 8;
__global__ void __launch_bounds__(kPsdWarps * 32) ps_sideinfo_kernel(int16_t *records, long long n) {
  __shared__ __align__(16) int16_t ws[kPsdWarps][XAAC_PSD_WORDS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int16_t *r = ws[warp];
  const long long warps_total = (long long)gridDim.x * kPsdWarps;
  for (long long u = (long long)blockIdx.x * kPsdWarps + warp; u < n; u += warps_total) {
    int4 *g = reinterpret_cast<int4 *>(records + u * XAAC_PSD_WORDS);
    int4 *s4 = reinterpret_cast<int4 *>(r);
    __syncwarp();
    for (int i = lane; i < XAAC_PSD_WORDS / 8; i += 32) s4[i] = g[i];
    __syncwarp();
    const int iid_mode = min(max((int)r[XAAC_PSD_IID_MODE], 0), 2), icc_mode = min(max((int)r[XAAC_PSD_ICC_MODE], 0), 2);
    const int max_cols = r[XAAC_PSD_FRAME_SIZE] == 960 ? 30 : 32;
    int num_env = r[XAAC_PSD_DATA_PRESENT] ? min(max((int)r[XAAC_PSD_NUM_ENV], 0), 5) : 0;
    if (lane < 2) {  // lane 0: IID, lane 1: ICC
      int16_t *tab = r + (lane ? XAAC_PSD_ICC_TABLE : XAAC_PSD_IID_TABLE);
      const int16_t *prev0 = r + (lane ? XAAC_PSD_ICC_PREV : XAAC_PSD_IID_PREV);
      const bool enable = r[lane ? XAAC_PSD_ENABLE_ICC : XAAC_PSD_ENABLE_IID] != 0;
      const int mode = lane ? icc_mode : iid_mode;
      const i32 lv = r[XAAC_PSD_IID_QUANT] ? 15 : 7;
      for (int e = 0; e < num_env; e++)
        ps_decode_set(tab + 34 * e, e ? tab + 34 * (e - 1) : prev0, enable, r[(lane ? XAAC_PSD_ICC_DT : XAAC_PSD_IID_DT) + e] != 0, mode,
                      lane ? 0 : -lv, lane ? 7 : lv);
    }
    __syncwarp();
    if (num_env == 0) {  // no PS data in this frame: one envelope holding the previous parameters (ps_bitdec.c:199-215)
      num_env = 1;
      for (int i = lane; i < 68; i += 32) {
        const int set = i >= 34, b = set ? i - 34 : i;
        const bool en = r[set ? XAAC_PSD_ENABLE_ICC : XAAC_PSD_ENABLE_IID] != 0;
        r[(set ? XAAC_PSD_ICC_TABLE : XAAC_PSD_IID_TABLE) + b] = en ? r[(set ? XAAC_PSD_ICC_PREV : XAAC_PSD_IID_PREV) + b] : (int16_t)0;
      }
      __syncwarp();
    }
    for (int i = lane; i < 68; i += 32) {  // the last envelope becomes the previous one (:217-223)
      const int set = i >= 34, b = set ? i - 34 : i;
      r[(set ? XAAC_PSD_ICC_PREV : XAAC_PSD_IID_PREV) + b] = r[(set ? XAAC_PSD_ICC_TABLE : XAAC_PSD_IID_TABLE) + 34 * (num_env - 1) + b];
    }
    __syncwarp();
    int16_t *bp = r + XAAC_PSD_BORDER;
    if (r[XAAC_PSD_FRAME_CLASS] == 0) {  // fixed borders (:227-247)
      if (lane == 0) {
        const int shift = num_env == 2 ? 1 : (num_env == 4 ? 2 : 0);
        bp[0] = 0;
        for (int e = 1; e < num_env; e++) bp[e] = (int16_t)((e * max_cols) >> shift);
        bp[num_env] = (int16_t)max_cols;
      }
    } else {  // variable borders (:248-277)
      const bool extend = bp[num_env] < max_cols;
      __syncwarp();
      if (extend) {
        num_env++;
        for (int i = lane; i < 68; i += 32) {
          const int set = i >= 34, b = set ? i - 34 : i;
          int16_t *t = r + (set ? XAAC_PSD_ICC_TABLE : XAAC_PSD_IID_TABLE);
          t[34 * (num_env - 1) + b] = t[34 * (num_env - 2) + b];
        }
      }
      __syncwarp();
      if (lane == 0) {
        bp[0] = 0;
        if (extend) bp[num_env] = (int16_t)max_cols;
        for (int e = 1; e < num_env; e++) {
          int thr = max_cols - (num_env - e);
          if (bp[e] > thr) {
            bp[e] = (int16_t)thr;
          } else {
            thr = bp[e - 1] + 1;
            if (bp[e] < thr) bp[e] = (int16_t)thr;
          }
        }
      }
    }
    __syncwarp();
    // 34 -> 20 bands, envelope per lane (:279-285)
    if (lane < num_env && iid_mode == 2) ps_map_34_to_20(r + XAAC_PSD_IID_TABLE + 34 * lane);
    if (lane >= 8 && lane - 8 < num_env && icc_mode == 2) ps_map_34_to_20(r + XAAC_PSD_ICC_TABLE + 34 * (lane - 8));
    if (lane == 31) {
      r[XAAC_PSD_NUM_ENV] = (int16_t)num_env;
      r[XAAC_PSD_DATA_PRESENT] = 0;
    }
    __syncwarp();
    for (int i = lane; i < XAAC_PSD_WORDS / 8; i += 32) g[i] = s4[i];
  }
}

cudaError_t launch_ps_sideinfo(int16_t *records, long long n, int num_sms, cudaStream_t stream) {
  long long need = (n + kPsdWarps - 1) / kPsdWarps;
  long long grid = (long long)num_sms * 8;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  ps_sideinfo_kernel<<<(unsigned)grid, kPsdWarps * 32, 0, stream>>>(records, n);
  return cudaGetLastError();
}

cudaError_t launch_sbr_sideinfo(int16_t *records, long long n, const uint8_t *misc_rom, int num_sms, cudaStream_t stream) {
  long long need = (n + kSdWarps - 1) / kSdWarps;
  long long grid = (long long)num_sms * 8;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  sbr_sideinfo_kernel<<<(unsigned)grid, kSdWarps * 32, 0, stream>>>(records, n, misc_rom);
  return cudaGetLastError();
}

}  // namespace xb
