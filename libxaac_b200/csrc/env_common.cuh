// env_common.cuh — (mantissa, exponent) helpers and table copies shared by the envelope-adjuster kernels
// (envcalc_kernel.cu: complex HQ path; sbr_lp_kernel.cu: low-power real path).  Reference lines are cited per function.
#pragma once
#include <cstdint>
#include "fixmath.cuh"

namespace xb {

constexpr int kMaxB = 56;  // MAX_FREQ_COEFFS

struct EnvRomS {            // block-shared copy of the small tables
  int16_t lim_gains[8];
  int16_t smooth[4];
  int16_t inv_int[50];
  int16_t inv_table[256];
  int16_t sqrt_table[258];
};

XB_DEV i32 mult16_shl_sat_(i32 a, i32 b) { return sat16((a * b) >> 15); }
XB_DEV i32 mult16_shl_(i32 a, i32 b) { return sext16((a * b) >> 15); }
XB_DEV i32 mult16_(i32 a, i32 b) { return sext16((a * b) >> 16); }
XB_DEV i32 shr32_dir_sat_limit(i32 a, int b) {  // ops.h:104
  if (b < 0) return shl32_sat(a, -b);
  return a >> (b > 31 ? 31 : b);
}
XB_DEV i32 shr32_dir(i32 a, int b) { return b < 0 ? shl32(a, -b) : shr32(a, b); }

// basic_funcs.c:66-99
XB_DEV int mant_div(i32 a, i32 b, i32 &res, const EnvRomS &r) {
  const int pre = norm32(b) - 16;
  int post;
  const int idx = (lsl(b, pre) >> 5) & 0x1ff;
  if (idx == 0) {
    post = norm32(a) - 16;
    res = sext16(lsl(a, post));
  } else {
    const i32 ratio = (i32)r.inv_table[(idx - 1) >> 1] * a;
    post = norm32(ratio) - 1;
    res = sext16(lsl(ratio, post) >> 15);
  }
  return pre - post;
}

// basic_funcs.c:101-128
XB_DEV void mant_exp_sqrt(int16_t *v, const EnvRomS &r) {
  i32 m = v[0], e = v[1];
  if (m > 0) {
    const int pre = norm32(m) - 16;
    e -= pre;
    const int idx = (lsl(m, pre) >> 5) & 0x1ff;
    i32 res = r.sqrt_table[idx >> 1];
    if (e & 1) {
      res = (res * 0x5a82) >> 16;
      e += 3;
    }
    v[0] = (int16_t)res;
    v[1] = (int16_t)(e >> 1);
  } else {
    v[0] = 0;
    v[1] = -16;
  }
}

// the (mantissa, exponent) running sum the reference open-codes in avggain_calc / noiselimiting (env_calc.c:1493-1527,
// 326-336): the operand with the smaller exponent is shifted down (ixheaac_shr32: count & 0xff, >= 31 gives the sign) and
// added.  Branch-free, because the lanes of a warp walk different limiter bands here.
XB_DEV void acc_add(i32 &am, i32 &ae, i32 m, i32 e) {
  const i32 d = e - ae;
  const bool ge = d >= 0;
  const int s = min((ge ? d : -d) & 0xff, 31);
  const i32 big = ge ? m : am, small = ge ? am : m;
  am = big + (small >> s);
  ae = ge ? e : ae;
}

// env_calc.c:1382-1452
XB_DEV void subbandgain(i32 ref_m, i32 noise_m, i32 est_m, i32 est_e, i32 noise_e, i32 ref_e, bool present, bool mapped,
                        bool noise_absc, int16_t *gain, int16_t *noise, int16_t *sine, const EnvRomS &r) {
  i32 v1m, v1e, v2m, v2e, v3m, v3e, q;
  if (est_m == 0) {
    est_m = 0x4000;
    est_e = 1;
  }
  v1m = mult16_shl_sat_(ref_m, noise_m);
  v1e = sext16(ref_e + noise_e);
  {
    i32 accu, d = noise_e - 1;
    if (d >= 0) {
      accu = noise_m + shr32(0x4000, d);
      v2e = noise_e;
    } else {
      accu = shr32(noise_m, -d) + 0x4000;
      v2e = 1;
    }
    if ((accu < 0 ? -accu : accu) >= 0x8000) {
      accu >>= 1;
      v2e++;
    }
    v2m = sext16(accu);
  }
  int t = mant_div(v1m, v2m, q, r);
  noise[0] = (int16_t)q;
  noise[1] = (int16_t)(t + (v1e - v2e) + 1);
  if (present || !noise_absc) {
    v3m = mult16_shl_sat_(v2m, est_m);
    v3e = sext16(v2e + est_e);
  } else {
    v3m = est_m;
    v3e = est_e;
  }
  if (!present) {
    v1m = ref_m;
    v1e = ref_e;
  }
  t = mant_div(v1m, v3m, q, r);
  gain[0] = (int16_t)q;
  gain[1] = (int16_t)(t + (v1e - v3e) + 1);
  if (present && mapped) {
    t = mant_div(ref_m, v2m, q, r);
    sine[0] = (int16_t)q;
    sine[1] = (int16_t)(t + (ref_e - v2e) + 1);
  }
}

}  // namespace xb
