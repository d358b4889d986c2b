/*
 * oracle/src/envcalc.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain-C restatement of the fixed-point complex ("HQ", low_pow_flag = 0) SBR envelope adjuster of libxaac
 * (SURVEY.md §8a-C): ixheaacd_calc_sbrenvelope and everything below it.  Cites reference lines (paths relative
 * to /root/reference).  Pinned against records tapped from real decodes of the compiled reference
 * (tests/golden/envcalc_tapped.npz) and against the compiled reference driven with perturbed copies of them.
 *
 * All (mantissa, exponent) quantities are WORD16 pairs as in the reference; arrays are indexed by band relative
 * to max_qmf_subband_aac ("c" in the reference) unless stated.
 */
#include <string.h>
#include "fixmath.h"
#include "xaac_oracle.h"

#define MASK_M 0xffc0     /* decoder/ixheaacd_env_extr.h:27 */
#define MASK_E 0x3f       /* :28 */
#define NRG_EXP_OFFSET 16 /* :32 */
#define NOISE_EXP_OFFSET 38
#define MAX_GAIN_EXP 34 /* decoder/ixheaacd_sbrdecsettings.h:63 */
#define MAXB 56         /* MAX_FREQ_COEFFS */

typedef struct {
  const i16 *lim_gains, *smooth_filter, *inv_int;
  const i32 *rand_ph;
  const i16 *inv_table, *sqrt_table;
} env_rom_t;

/* The reference shifts promoted WORD16 values by run-time counts that may exceed 31 here (`x >> diff`).  That is
 * undefined in ISO C; the reference build this oracle is pinned to (gcc, x86-64) executes SAR/SHL, which use the count
 * modulo 32.  The oracle states that behaviour explicitly. */
#define X86_SHIFT(c) ((c) & 31)

typedef struct { i16 m, e; } me_t;

/* decoder/ixheaacd_basic_funcs.c:66-99 */
static int mant_div(i16 a, i16 b, i16 *res, const env_rom_t *r) {
  int pre = ox_norm32(b) - 16, post;
  int idx = ((i32)((u32)(i32)b << pre) >> 5) & 0x1ff;
  if (idx == 0) {
    post = ox_norm32(a) - 16;
    *res = (i16)((u32)(i32)a << post);
  } else {
    i32 ratio = (i32)r->inv_table[(idx - 1) >> 1] * (i32)a;
    post = ox_norm32(ratio) - 1;
    *res = (i16)((i32)((u32)ratio << post) >> 15);
  }
  return pre - post;
}

/* decoder/ixheaacd_basic_funcs.c:101-128 */
static void mant_exp_sqrt(me_t *v, const env_rom_t *r) {
  i32 m = v->m, e = v->e;
  if (m > 0) {
    int pre = ox_norm32((i16)m) - 16;
    e -= pre;
    int idx = ((i32)((u32)m << pre) >> 5) & 0x1ff;
    i32 res = r->sqrt_table[idx >> 1];
    if (e & 1) {
      res = (res * 0x5a82) >> 16;
      e += 3;
    }
    v->m = (i16)res;
    v->e = (i16)(e >> 1);
  } else {
    v->m = 0;
    v->e = -16;
  }
}

/* the (mant, exp) running sum the reference open-codes in avggain_calc / noiselimiting
 * (decoder/ixheaacd_env_calc.c:1493-1527, 326-336) */
static void acc_add(i32 *am, i32 *ae, i32 m, i32 e) {
  i32 d = e - *ae;
  if (d >= 0) {
    *am = m + ox_shr32(*am, d);
    *ae = e;
  } else {
    *am = ox_shr32(m, -d) + *am;
  }
}

/* decoder/ixheaacd_env_calc.c:1454-1562 with flag == 0 (the noise limiter's use) */
static void avggain(const me_t *orig, const me_t *est, int b0, int b1, i16 *sum_m, i16 *sum_e, i16 *gain_m, i16 *gain_e,
                    const env_rom_t *r) {
  i32 om = 0, oe = 0, em = 0, ee = 0;
  for (int k = b0; k < b1; k++) {
    acc_add(&om, &oe, orig[k].m, orig[k].e);
    acc_add(&em, &ee, est[k].m, est[k].e);
  }
  int nv = 16 - ox_pnorm32(om);
  if (nv > 0) { om >>= nv; oe += nv; }
  nv = 16 - ox_pnorm32(em);
  if (nv > 0) { em >>= nv; ee += nv; }
  int t = mant_div((i16)om, (i16)em, gain_m, r);
  *gain_e = (i16)(t + ((i16)oe - (i16)ee) + 1);
  *sum_m = (i16)om;
  *sum_e = (i16)oe;
}

/* decoder/ixheaacd_env_calc.c:1382-1452 */
static void subbandgain(i16 ref_m, i16 noise_m, i16 est_m, i16 est_e, i16 noise_e, i16 ref_e, int sine_present,
                        int sine_mapped, int noise_absc, me_t *gain, me_t *noise, me_t *sine, const env_rom_t *r) {
  i16 v1m, v1e, v2m, v2e, v3m, v3e;
  if (est_m == 0) { est_m = 0x4000; est_e = 1; }
  v1m = ox_mult16_shl_sat(ref_m, noise_m);
  v1e = (i16)(ref_e + noise_e);
  {
    i32 accu, d = noise_e - 1;
    if (d >= 0) { accu = noise_m + ox_shr32(0x4000, d); v2e = noise_e; }
    else { accu = ox_shr32((i32)noise_m, -d) + 0x4000; v2e = 1; }
    if ((accu < 0 ? -accu : accu) >= 0x8000) { accu >>= 1; v2e++; }
    v2m = (i16)accu;
  }
  noise->e = (i16)(mant_div(v1m, v2m, &noise->m, r) + (v1e - v2e) + 1);
  if (sine_present || !noise_absc) { v3m = ox_mult16_shl_sat(v2m, est_m); v3e = (i16)(v2e + est_e); }
  else { v3m = est_m; v3e = est_e; }
  if (!sine_present) { v1m = ref_m; v1e = ref_e; }
  gain->e = (i16)(mant_div(v1m, v3m, &gain->m, r) + (v1e - v3e) + 1);
  if (sine_present && sine_mapped) sine->e = (i16)(mant_div(ref_m, v2m, &sine->m, r) + (ref_e - v2e) + 1);
}

/* decoder/ixheaacd_env_calc.c:1211-1296, low_pow_flag == 0.  m: matrix row 0, [slot][re 64 | im 64]. */
static void energy_per_subband(const i32 *m, int start, int next, int b0, int b1, int frame_exp, me_t *est,
                               const env_rom_t *r) {
  i16 inv_width = r->inv_int[next - start];
  frame_exp <<= 1;
  for (int k = b0; k < b1; k++, est++) {
    i32 max_val = 1;
    for (int l = start; l < next; l++) {
      i32 a = ox_abs_nrm(m[128 * l + k]), b = ox_abs_nrm(m[128 * l + 64 + k]);
      if (a > max_val) max_val = a;
      if (b > max_val) max_val = b;
    }
    int pre = ox_pnorm32(max_val) - 4, shift = 16 - pre;
    i32 accu = 0;
    for (int l = start; l < next; l++)
      for (int c = 0; c < 2; c++) {
        i32 v = m[128 * l + 64 * c + k];
        i16 t = shift > 0 ? (i16)(v >> shift) : (i16)((u32)v << -shift);
        accu = ox_add(accu, (i32)t * t);
      }
    if (accu != 0) {
      shift = -ox_pnorm32(accu);
      i16 sum_m = (i16)ox_shr32_dir_sat_limit(accu, 16 + shift);
      est->m = ox_mult16_shl_sat(sum_m, inv_width);
      shift -= pre << 1;
      est->e = (i16)(frame_exp + shift + 1);
    } else {
      est->m = est->e = 0;
    }
  }
}

/* decoder/ixheaacd_env_calc.c:1159-1207 (complex branch): headroom of a [slot range] x [band range] block */
int xo_expsubbandsamples_hq(const i32 *m, int b0, int b1, int s0, int s1) {
  i32 mx = 1;
  for (int l = s0; l < s1; l++)
    for (int k = b0; k < b1; k++) mx |= ox_abs_nrm(m[128 * l + k]) | ox_abs_nrm(m[128 * l + 64 + k]);
  return (i16)ox_pnorm32(mx);
}

/* decoder/ixheaacd_env_calc.c:1099-1157 (complex branch) */
void xo_adjust_scale_hq(i32 *m, int b0, int b1, int s0, int s1, int shift) {
  if (shift == 0) return;
  if (shift > 31) shift = 31;
  if (shift < -31) shift = -31;
  for (int l = s0; l < s1; l++)
    for (int k = b0; k < b1; k++)
      for (int c = 0; c < 2; c++) {
        i32 *p = &m[128 * l + 64 * c + k];
        *p = shift > 0 ? (i32)((u32)*p << shift) : (*p >> -shift);
      }
}

/* decoder/ixheaacd_env_calc.c:1298-1380, low_pow_flag == 0 */
static void energy_per_sfb(const i32 *m, int num_sfb, const i16 *tbl, int start, int next, int max_qmf, int frame_exp,
                           me_t *est, const env_rom_t *r) {
  i16 inv_width = r->inv_int[next - start];
  frame_exp <<= 1;
  for (int j = 0; j < num_sfb; j++) {
    int li = tbl[j], ui = tbl[j + 1];
    if (li < max_qmf) continue;
    int pre = xo_expsubbandsamples_hq(m, li, ui, start, next) - 4;
    i32 accumulate = 0;
    for (int k = li; k < ui; k++) {
      int s = 16 - pre;
      i32 line = 0;
      if (s > 31) s = 31;
      for (int l = start; l < next; l++)
        for (int c = 0; c < 2; c++) {
          i16 t = (i16)ox_shr32_dir(m[128 * l + 64 * c + k], s);
          line = ox_add_sat(line, (i32)t * t);
        }
      accumulate = ox_add_sat(accumulate, ox_shr32(line, 9));
    }
    int shift = ox_pnorm32(accumulate);
    i16 sum_m = (i16)ox_shr32_dir_sat_limit(accumulate, 16 - shift);
    i32 sum_e = 0;
    if (sum_m != 0) {
      sum_m = ox_mult16_shl_sat(sum_m, inv_width);
      sum_m = ox_mult16_shl_sat(sum_m, r->inv_int[ui - li]);
      sum_e = (frame_exp + 10) - shift - (pre << 1);
    }
    for (int k = li; k < ui; k++, est++) { est->m = sum_m; est->e = (i16)sum_e; }
  }
}

/* decoder/ixheaacd_env_calc.c:229-421 */
static void noise_limiting(const i16 *prm, int skip, const me_t *orig, const me_t *est, me_t *gain, me_t *noise,
                           me_t *sine, const i16 *lim_gain, int noise_absc, const env_rom_t *r) {
  const i16 *lim = prm + XO_ENV_LIM_TBL;
  for (int c = 0; c < prm[XO_ENV_NUM_LF_BANDS]; c++) {
    int b0 = lim[c] > skip ? lim[c] - skip : 0, b1 = lim[c + 1] > skip ? lim[c + 1] - skip : 0;
    if (b0 >= b1) continue;
    i16 sum_m, sum_e, mg_m, mg_e;
    avggain(orig, est, b0, b1, &sum_m, &sum_e, &mg_m, &mg_e, r);
    i32 mt = ox_shl32((i32)mg_m * lim_gain[0], 1);
    mg_e = (i16)(mg_e + lim_gain[1]);
    int nv = ox_norm32(mt);
    mg_e = (i16)(mg_e - nv);
    mg_m = (i16)((i32)((u32)mt << nv) >> 16);
    if (mg_e >= MAX_GAIN_EXP) { mg_m = 0x3000; mg_e = MAX_GAIN_EXP; }
    for (int k = b0; k < b1; k++)
      if (gain[k].e > mg_e || (gain[k].e == mg_e && gain[k].m > mg_m)) {
        i16 na_m;
        i16 na_e = (i16)mant_div(mg_m, gain[k].m, &na_m, r);
        na_e = (i16)(na_e + (mg_e - gain[k].e) + 1);
        noise[k].m = (i16)(ox_shl32_dir_sat_limit(ox_shl32((i32)noise[k].m * na_m, 1), na_e) >> 16);
        gain[k].m = mg_m;
        gain[k].e = mg_e;
      }
    i32 am = 0, ae = 0;
    for (int k = b0; k < b1; k++) {
      acc_add(&am, &ae, ((i32)gain[k].m * est[k].m) >> 15, gain[k].e + est[k].e);
      if (sine[k].m != 0) acc_add(&am, &ae, sine[k].m, sine[k].e);
      else if (!noise_absc) acc_add(&am, &ae, noise[k].m, noise[k].e);
    }
    nv = 16 - ox_norm32(am);
    if (nv > 0) { am >>= nv; ae += nv; }
    i16 bg_m;
    i16 bg_e = (i16)mant_div(sum_m, (i16)am, &bg_m, r);
    bg_e = (i16)(bg_e + (sum_e - (i16)ae) + 1);
    if (bg_e > 2 || (bg_e == 2 && bg_m > 0x5061)) { bg_m = 0x5061; bg_e = 2; }
    for (int k = b0; k < b1; k++) {
      gain[k].m = ox_mult16_shl(gain[k].m, bg_m);
      sine[k].m = ox_mult16_shl(sine[k].m, bg_m);
      noise[k].m = ox_mult16_shl(noise[k].m, bg_m);
      gain[k].e = (i16)(gain[k].e + bg_e);
      sine[k].e = (i16)(sine[k].e + bg_e);
      noise[k].e = (i16)(noise[k].e + bg_e);
    }
  }
}



/* decoder/ixheaacd_env_calc.c:1080-1097 */
static void noise_rescale(i16 *p, int diff, int n, int stride) {
  if (diff > 0) for (int k = 0; k < n; k++) p[k * stride] = (i16)(p[k * stride] >> X86_SHIFT(diff));
  else if (diff < 0) for (int k = 0; k < n; k++) p[k * stride] = (i16)((u32)(i32)p[k * stride] << X86_SHIFT(-diff));
}

/* per-stream envelope-adjuster state record (WORD16[XO_ENV_ST_WORDS]) = ia_sbr_calc_env_struct
 * (decoder/ixheaacd_env_calc.h:24-33) */
typedef struct {
  i16 filt_me[2 * MAXB], filt_noise[MAXB], filt_noise_e, start_up, ph_index, trans_prev, harm_index, harm_prev[MAXB];
} env_state_t;

/* decoder/ixheaacd_env_dec.c:845-923 with ixheaacd_harm_idx_zerotwo / _onethree
 * (decoder/ixheaacd_env_calc.c:1759-1898).  re/im point at band sub_band_start (= max_qmf_subband_aac). */
static void adj_timeslot(i32 *re, i32 *im, i16 *filt_me, i16 *filt_noise, const me_t *gain, const me_t *noise,
                         const me_t *sine, i16 noise_e, env_state_t *st, int sb_start, int nb, i16 scale_change,
                         i16 smooth_ratio, int noise_absc, const env_rom_t *r) {
  i16 direct = ox_sub16_sat(0x7fff, smooth_ratio);
  int index = st->ph_index, harm = st->harm_index, finv = sb_start & 1;
  scale_change = (i16)(scale_change - 1);
  const i32 *rnd = r->rand_ph + index;
  st->ph_index = (i16)((index + nb) & 511);
  if (smooth_ratio)
    for (int k = 0; k < nb; k++) {
      i16 t = (i16)(ox_mult16(smooth_ratio, filt_me[2 * k]) + ox_mult16(direct, gain[k].m));
      i16 t1 = (i16)(ox_mult16(smooth_ratio, filt_noise[k]) + ox_mult16(direct, noise[k].m));
      filt_me[2 * k] = (i16)(t << 1);
      filt_noise[k] = (i16)(t1 << 1);
    }
  if (harm == 1) finv = !finv;
  for (int k = 0; k < nb; k++) {
    i16 g = smooth_ratio ? filt_me[2 * k] : gain[k].m;
    i16 nz = smooth_ratio ? filt_noise[k] : noise[k].m;
    i32 sr = ox_mul32x16(re[k], g), si = ox_mul32x16(im[k], g);
    int shift = ox_sub16(gain[k].e, scale_change);
    if (shift > 0) { sr = ox_shl32(sr, shift); si = ox_shl32(si, shift); }
    else { sr = ox_shr32(sr, -shift); si = ox_shr32(si, -shift); }
    if (sine[k].m != 0) {
      int t = ox_sub16(sine[k].e, noise_e);
      i32 sl;
      if (t > 0) sl = ox_shl32(sine[k].m, t);
      else if (harm & 1) sl = ox_shr32(sine[k].m, -t);
      else sl = ox_shr32(sine[k].m, t); /* :1796 passes the non-positive count unnegated; shr32 masks it to 8 bits */
      if (harm == 0) sr = ox_add_sat(sr, sl);
      else if (harm == 2) sr = ox_sub_sat(sr, sl);
      else si = finv ? ox_add_sat(si, sl) : ox_sub_sat(si, sl);
    } else if (!noise_absc) {
      i32 rv = rnd[k + 1];
      i32 pr = (i32)(i16)(rv >> 16) * nz, pi = (i32)(i16)rv * nz;
      sr = ox_add_sat(sr, pr == 0x40000000 ? OX_MAX32 : ox_shl32(pr, 1));
      si = ox_add_sat(si, pi == 0x40000000 ? OX_MAX32 : ox_shl32(pi, 1));
    }
    re[k] = sr;
    im[k] = si;
    finv = !finv;
  }
  st->harm_index = (i16)((harm + 1) & 3);
}

#include "envcalc_lp.inc"

/* decoder/ixheaacd_env_calc.c:692-1015 for non-ELD/LD object types, 1024-sample core frames (num_time_slots 16,
 * max_cols 32).  low_pow = 0: complex matrix, 128 words per slot; low_pow = 1: real matrix, 64 words per slot, `deg` =
 * degree_alias[64] from the low-power HF generator.  Returns 0 or 0x80000000 (IA_FATAL_ERROR) like the reference. */
static int calc_sbrenvelope(const uint8_t *env_rom, const uint8_t *misc_rom, const i16 *prm, i16 *sf, i16 *state,
                            i32 *matrix, int low_pow, const i16 *deg) {
  env_rom_t rom;
  rom.lim_gains = (const i16 *)(env_rom + XO_EROM_LIM_GAINS);
  rom.smooth_filter = (const i16 *)(env_rom + XO_EROM_SMOOTH);
  rom.inv_int = (const i16 *)(env_rom + XO_EROM_INV_INT);
  rom.rand_ph = (const i32 *)(env_rom + XO_EROM_RAND_PH);
  rom.inv_table = (const i16 *)(misc_rom + XO_MROM_INV_TABLE);
  rom.sqrt_table = (const i16 *)(misc_rom + XO_MROM_SQRT_TABLE);
  env_state_t *st = (env_state_t *)state;

  const int num_env = prm[XO_ENV_NUM_ENV], trans_env = prm[XO_ENV_TRANSIENT_ENV];
  const i16 *border = prm + XO_ENV_BORDER_VEC, *freq_res = prm + XO_ENV_FREQ_RES;
  const i16 *nborder = prm + XO_ENV_NOISE_BORDER_VEC;
  const int num_sf[2] = {prm[XO_ENV_NUM_SF_LO], prm[XO_ENV_NUM_SF_HI]};
  const i16 *ftab[2] = {prm + XO_ENV_FREQ_LO, prm + XO_ENV_FREQ_HI};
  const i16 *fnoise = prm + XO_ENV_FREQ_NOISE;
  const int num_nf = prm[XO_ENV_NUM_NF_BANDS];
  const int sb_start = prm[XO_ENV_SUB_BAND_START], sb_end = prm[XO_ENV_SUB_BAND_END];
  const int max_qmf = prm[XO_ENV_MAX_QMF_SUBBAND], max_qmf_prev = prm[XO_ENV_MAX_QMF_SUBBAND_PREV];
  const int num_sub_bands = sb_end - sb_start, skip = max_qmf - sb_start, bands = num_sub_bands - skip;
  const i16 *noise_floor = prm + XO_ENV_NOISE_FLOOR;
  const i16 *sf_arr = prm + XO_ENV_SF_ARR;
  int8_t sine_mapped[MAXB], alias_red[64 + MAXB];
  memset(alias_red, 0, sizeof(alias_red));
  me_t est[MAXB], gain[MAXB], noise[MAXB], sine[MAXB], orig[MAXB];
  memset(est, 0, sizeof(est)); memset(gain, 0, sizeof(gain)); memset(noise, 0, sizeof(noise));
  memset(sine, 0, sizeof(sine)); memset(orig, 0, sizeof(orig));

  /* decoder/ixheaacd_sbrdec_lpfuncs.c:529-560 */
  memset(sine_mapped, 8, sizeof(sine_mapped));
  for (int i = num_sf[1] - 1, p = 0; i >= 0; i--, p++) {
    int old = st->harm_prev[p];
    int add = prm[XO_ENV_ADD_HARMONICS + i];
    st->harm_prev[p] = (i16)(int8_t)add;
    if (add) {
      int q = ((ftab[1][i + 1] + ftab[1][i]) - (ftab[1][0] << 1)) >> 1;
      sine_mapped[q] = old ? 0 : (int8_t)trans_env;
    }
  }

  int adj_e, final_e = 0;
  { /* :772-791 */
    int first = (max_qmf_prev > max_qmf ? max_qmf_prev : max_qmf) - sb_start;
    i16 mx = 0;
    for (int i = first; i < num_sub_bands; i++) if (st->filt_noise[i] > mx) mx = st->filt_noise[i];
    adj_e = (st->filt_noise_e - ox_norm32(mx)) - 16;
  }
  { /* :793-841 */
    const i16 *p = sf_arr;
    for (int i = 0; i < num_env; i++) {
      int mx = NRG_EXP_OFFSET - 16;
      for (int j = 0; j < num_sf[freq_res[i]]; j++) { int t = *p++ & MASK_E; if (t > mx) mx = t; }
      int t = ((mx - NRG_EXP_OFFSET) + 13) >> 1;
      if (border[i] < 16 && t > adj_e) adj_e = (i16)t;
      if (border[i + 1] > 16 && t > final_e) final_e = (i16)t;
    }
  }

  int m = 0, nf_idx = 0;
  for (int i = 0; i < num_env; i++) {
    int start = 2 * border[i], end = 2 * border[i + 1], fr = freq_res[i];
    if (start >= 38 || end > 38) return (int)0x80000000;
    if (nf_idx >= 2) return (int)0x80000000;
    if (border[i] == nborder[nf_idx + 1]) { noise_floor += num_nf; nf_idx++; }
    int noise_absc = (i == trans_env || i == st->trans_prev);
    int smooth_len = noise_absc ? 0 : ((1 - prm[XO_ENV_SMOOTHING_MODE]) << 2);
    int input_e = 15 - sf[XO_SF_HB];
    if (low_pow) {
      if (prm[XO_ENV_INTERPOL_FREQ]) energy_per_subband_lp(matrix, start, end, max_qmf, sb_end, input_e, est, &rom);
      else energy_per_sfb_lp(matrix, num_sf[fr], ftab[fr], start, end, max_qmf, input_e, est, &rom);
    } else if (prm[XO_ENV_INTERPOL_FREQ]) energy_per_subband(matrix, start, end, max_qmf, sb_end, input_e, est, &rom);
    else energy_per_sfb(matrix, num_sf[fr], ftab[fr], start, end, max_qmf, input_e, est, &rom);
    if (ftab[fr][0] < sb_start) return (int)0x80000000;

    { /* decoder/ixheaacd_env_calc.c:616-688 */
      int ui_noise = fnoise[1], nb = 0, c = 0, sm = 0, ar = ftab[fr][0] - sb_start;
      i16 nm = (i16)(noise_floor[0] & MASK_M), ne = (i16)((noise_floor[0] & MASK_E) - NOISE_EXP_OFFSET);
      for (int j = 0; j < num_sf[fr]; j++) {
        int li = ftab[fr][j], ui = ftab[fr][j + 1];
        i16 v = sf_arr[m + j];
        i16 ref_e = (i16)((v & MASK_E) - NRG_EXP_OFFSET), ref_m = (i16)(v & MASK_M);
        int present = 0;
        for (int k = li; k < ui; k++) if (i >= sine_mapped[sm++]) present = 1;
        for (int k = li; k < ui; k++) {
          alias_red[ar++] = (int8_t)!present;
          if (k >= ui_noise) {
            nb++;
            ui_noise = fnoise[nb + 1];
            nm = (i16)(noise_floor[nb] & MASK_M);
            ne = (i16)((noise_floor[nb] & MASK_E) - NOISE_EXP_OFFSET);
          }
          if (k >= max_qmf) {
            orig[c].m = ref_m; orig[c].e = ref_e;
            sine[c].m = sine[c].e = 0;
            subbandgain(ref_m, nm, est[c].m, est[c].e, ne, ref_e, present, i >= sine_mapped[skip + c], noise_absc,
                        &gain[c], &noise[c], &sine[c], &rom);
            c++;
          }
        }
      }
    }
    m += num_sf[fr];
    noise_limiting(prm, skip, orig, est, gain, noise, sine, rom.lim_gains + 2 * prm[XO_ENV_LIMITER_GAINS], noise_absc,
                   &rom);
    if (low_pow) alias_reduction(deg + sb_start, gain, est, alias_red, num_sub_bands, &rom);
    i16 noise_e = (i16)(start < 32 ? adj_e : final_e);
    if (low_pow) conv_erg_to_amplitude_lp(bands, noise_e, sine, gain, noise, &rom);
    else for (int k = 0; k < bands; k++) { /* :450-477 */
      mant_exp_sqrt(&sine[k], &rom);
      mant_exp_sqrt(&gain[k], &rom);
      mant_exp_sqrt(&noise[k], &rom);
      int shift = (noise_e - noise[k].e) - 4;
      if (shift > 0) noise[k].m = (i16)(noise[k].m >> (shift > 31 ? 31 : shift));
      else noise[k].m = (i16)((i32)noise[k].m << (shift < -31 ? 31 : -shift));
    }

    /* decoder/ixheaacd_env_calc.c:479-614 */
    i16 *fme = st->filt_me + 2 * skip, *fno = st->filt_noise + skip;
    if (st->start_up) {
      st->start_up = 0;
      st->filt_noise_e = noise_e;
      for (int k = 0; k < bands; k++) { fme[2 * k] = gain[k].m; fme[2 * k + 1] = gain[k].e; fno[k] = noise[k].m; }
    } else { /* :1017-1058 */
      for (int k = 0; k < bands; k++) {
        i32 fe = fme[2 * k + 1], fm = fme[2 * k], diff = gain[k].e - fe;
        if (diff >= 0) {
          fme[2 * k + 1] = gain[k].e;
          fme[2 * k] = (i16)(fme[2 * k] >> X86_SHIFT(diff));
        } else {
          int reserve = ox_norm32(fm) - 16;
          if (diff + reserve >= 0) {
            fme[2 * k] = (i16)((u32)fm << -diff);
            fme[2 * k + 1] = (i16)(fe + diff);
          } else {
            fme[2 * k] = (i16)((u32)fm << reserve);
            fme[2 * k + 1] = (i16)(fe - reserve);
            int shift = -(reserve + diff);
            gain[k].m = (i16)(gain[k].m >> X86_SHIFT(shift));
            gain[k].e = (i16)(gain[k].e + shift);
          }
        }
      }
    }
    for (int l = start; l < end; l++) {
      int scale_change;
      if (l < 32) scale_change = adj_e - input_e;
      else {
        scale_change = final_e - input_e;
        if (l == 32 && start < 32) {
          int diff = final_e - noise_e;
          noise_e = (i16)final_e;
          noise_rescale(&noise[0].m, diff, bands, 2);
        }
      }
      noise_rescale(st->filt_noise, st->filt_noise_e - noise_e, num_sub_bands, 1);
      st->filt_noise_e = noise_e;
      if (low_pow) { /* :556-584 */
        int index = st->ph_index, harm = st->harm_index, finv = max_qmf & 1;
        i32 *re = matrix + 64 * l + max_qmf;
        const i32 *rnd = rom.rand_ph + index + 1;
        st->ph_index = (i16)((index + num_sub_bands) & 511);
        st->harm_index = (i16)((harm + 1) & 3);
        if (!(harm & 1)) {
          harm_idx_zerotwo_lp(re, gain, scale_change, sine, rnd, noise, num_sub_bands, noise_absc, harm);
        } else {
          int nz = (noise_e - 16) - (i16)(15 - sf[XO_SF_LB]);
          finv = !finv;
          finv = (finv << 1) - 1;
          if (harm == 3) finv = -finv;
          harm_idx_onethree_lp(re, gain, scale_change, sine, rnd, noise, num_sub_bands, noise_absc, finv, nz, max_qmf);
        }
        continue;
      }
      i16 ratio = (l - start) < smooth_len ? rom.smooth_filter[l - start] : 0;
      adj_timeslot(matrix + 128 * l + max_qmf, matrix + 128 * l + 64 + max_qmf, fme, fno, gain, noise, sine,
                   (i16)(noise_e - 16), st, max_qmf, bands, (i16)scale_change, ratio, noise_absc, &rom);
    }
    for (int k = 0; k < bands; k++) { fme[2 * k] = gain[k].m; fno[k] = noise[k].m; } /* :1060-1078 */
  }

  { /* :956-1007 */
    int first_start = border[0] * 2, ov_reserve = 0, reserve = 0;
    if (prm[XO_ENV_CHANNEL_MODE] == 3) {
      ov_reserve = low_pow ? xo_expsubbandsamples_lp(matrix, max_qmf, sb_end, 0, first_start)
                           : xo_expsubbandsamples_hq(matrix, max_qmf, sb_end, 0, first_start);
      reserve = low_pow ? xo_expsubbandsamples_lp(matrix, max_qmf, sb_end, first_start, 32)
                        : xo_expsubbandsamples_hq(matrix, max_qmf, sb_end, first_start, 32);
    }
    int ov_adj_e = 15 - sf[XO_SF_OV_HB];
    int output_e = (ov_adj_e - ov_reserve) > (adj_e - reserve) ? (ov_adj_e - ov_reserve) : (adj_e - reserve);
    void (*adj)(i32 *, int, int, int, int, int) = low_pow ? xo_adjust_scale_lp : xo_adjust_scale_hq;
    adj(matrix, max_qmf, sb_end, 0, first_start, ov_adj_e - output_e);
    adj(matrix, max_qmf, sb_end, first_start, prm[XO_ENV_NUM_TIME_SLOTS] * prm[XO_ENV_TIME_STEP], adj_e - output_e);
    sf[XO_SF_HB] = (i16)(15 - output_e);
    sf[XO_SF_OV_HB] = (i16)(15 - final_e);
  }
  st->trans_prev = (trans_env == num_env) ? 0 : -1;
  return 0;
}

int xo_calc_sbrenvelope_hq(const uint8_t *env_rom, const uint8_t *misc_rom, const i16 *prm, i16 *sf, i16 *state,
                           i32 *matrix) {
  return calc_sbrenvelope(env_rom, misc_rom, prm, sf, state, matrix, 0, 0);
}
int xo_calc_sbrenvelope_lp(const uint8_t *env_rom, const uint8_t *misc_rom, const i16 *prm, i16 *sf, i16 *state,
                           i32 *matrix, const i16 *degree_alias) {
  return calc_sbrenvelope(env_rom, misc_rom, prm, sf, state, matrix, 1, degree_alias);
}

void xo_calc_sbrenvelope_hq_batch(const uint8_t *env_rom, const uint8_t *misc_rom, const i16 *prm, i16 *sf, i16 *state,
                                  i32 *matrix, i32 *err, int n) {
  for (int u = 0; u < n; u++)
    err[u] = xo_calc_sbrenvelope_hq(env_rom, misc_rom, prm + (size_t)u * XO_ENV_PRM_WORDS, sf + (size_t)u * 8,
                                    state + (size_t)u * XO_ENV_ST_WORDS, matrix + (size_t)u * 38 * 128);
}
