/*
 * oracle/ref_shim.c — TEST INFRASTRUCTURE ONLY.
 *
 * Flat-array entry points around the UNMODIFIED reference decoder (ittiam-systems/libxaac), so that
 * tests (ctypes) and bench.py's cpu_baseline / --impl reference leg can call the reference's own stage
 * functions on plain buffers.  This file is OUR code; it is compiled against the reference headers
 * where they lie (/root/reference) and linked with the reference objects into oracle/_ref/libxaac_ref.so
 * by oracle/Makefile (target `ref`).  Nothing here is part of the product library.
 *
 * Every function states which reference function it drives (file:line in /root/reference).
 */
#include "ref_headers.h"

/* ------------------------------------------------------------------------------------------------
 * ROM access: lets tools/extract_rom.py and the tests read the reference's const tables.
 * ---------------------------------------------------------------------------------------------- */
const void *ref_rom_imdct_tables(int *bytes) {
  if (bytes) *bytes = (int)sizeof(ixheaacd_imdct_tables);
  return &ixheaacd_imdct_tables;
}

/* ------------------------------------------------------------------------------------------------
 * AAC IMDCT + window/OLA: drives ixheaacd_imdct_process (decoder/ixheaacd_lpfuncs.c:347) exactly the
 * way ixheaacd_aacdec_decodeframe does (decoder/ixheaacd_aacdecoder.c:988): 1024-sample frames,
 * AOT_AAC_LC, window tables wired as at aacdecoder.c:192-200.
 *   spec      [1024]  in, destroyed (used as FFT workspace by the reference)
 *   overlap   [512]   in/out
 *   prev_shape/prev_seq  in/out (ia_aac_dec_overlap_info.window_shape / window_sequence)
 *   out       1024 samples written at stride ch_fac
 *   returns qshift_adj
 * ---------------------------------------------------------------------------------------------- */
int ref_imdct_process(int32_t *spec, int32_t *overlap, int32_t *prev_shape, int32_t *prev_seq,
                      int32_t win_seq, int32_t win_shape, int32_t *out, int32_t ch_fac) {
  ia_aac_dec_overlap_info ovl;
  ia_ics_info_struct ics;
  ia_aac_dec_tables_struct tabs;
  WORD32 scratch[1024 + 64];
  memset(&ovl, 0, sizeof(ovl));
  memset(&ics, 0, sizeof(ics));
  memset(&tabs, 0, sizeof(tabs));
  tabs.pstr_imdct_tables = (ia_aac_dec_imdct_tables_struct *)&ixheaacd_imdct_tables;
  ovl.ptr_long_window[0] = ixheaacd_imdct_tables.only_long_window_sine;
  ovl.ptr_short_window[0] = ixheaacd_imdct_tables.only_short_window_sine;
  ovl.ptr_long_window[1] = ixheaacd_imdct_tables.only_long_window_kbd;
  ovl.ptr_short_window[1] = ixheaacd_imdct_tables.only_short_window_kbd;
  ovl.window_shape = (WORD16)*prev_shape;
  ovl.window_sequence = (WORD16)*prev_seq;
  ovl.ptr_overlap_buf = overlap;
  ics.window_shape = (WORD16)win_shape;
  ics.window_sequence = (WORD16)win_seq;
  ics.frame_length = 1024;
  ixheaacd_imdct_process(&ovl, spec, &ics, out, (WORD16)ch_fac, scratch, &tabs, AOT_AAC_LC, 0, 0);
  *prev_shape = ovl.window_shape;
  *prev_seq = ovl.window_sequence;
  return ics.qshift_adj;
}

/* Batch driver used by the CPU baseline: n units laid out unit-major, one private state per unit.
 * Returns nothing; qshift_adj[n] is filled. */
void ref_imdct_process_batch(int32_t *spec, int32_t *overlap, int32_t *prev_shape, int32_t *prev_seq,
                             const int32_t *win_seq, const int32_t *win_shape, int32_t *out,
                             int32_t *qshift_adj, int32_t n) {
  for (int32_t u = 0; u < n; u++) {
    qshift_adj[u] = ref_imdct_process(spec + (size_t)u * 1024, overlap + (size_t)u * 512, prev_shape + u,
                                      prev_seq + u, win_seq[u], win_shape[u], out + (size_t)u * 1024, 1);
  }
}

/* Leaf taps through the reference's function-selector pointers (decoder/ixheaacd_function_selector.h). */
int ref_calc_max_spectral_line(int32_t *spec, int32_t n) {
  return (*ixheaacd_calc_max_spectral_line)(spec, n);
}
int ref_inverse_transform(int32_t *spec, int32_t *scratch, int32_t expo, int32_t npoints) {
  return ixheaacd_inverse_transform(spec, scratch,
                                    (ia_aac_dec_imdct_tables_struct *)&ixheaacd_imdct_tables, expo, npoints);
}
void ref_post_twiddle(int32_t *out, int32_t *spec, int32_t npoints) {
  (*ixheaacd_post_twiddle)(out, spec, (ia_aac_dec_imdct_tables_struct *)&ixheaacd_imdct_tables, npoints);
}

/* ------------------------------------------------------------------------------------------------
 * Fixed-point SBR QMF banks.
 * ---------------------------------------------------------------------------------------------- */
const void *ref_rom_qmf_tables(int *bytes) {
  if (bytes) *bytes = (int)sizeof(ixheaacd_aac_qmf_dec_tables);
  return &ixheaacd_aac_qmf_dec_tables;
}
/* offsets of the members our QMF ROM blob relies on (checked by tests against the product's constants) */
int ref_rom_qmf_offsets(int *o) {
  int n = 0;
#define OFF(m) o[n++] = (int)offsetof(ia_qmf_dec_tables_struct, m)
  OFF(w_32); OFF(w_16); OFF(dig_rev_table2_32); OFF(dig_rev_table4_16); OFF(sbr_sin_cos_twiddle_l64);
  OFF(sbr_alt_sin_twiddle_l64); OFF(sbr_cos_sin_twiddle_ds_l32); OFF(sbr_sin_cos_twiddle_l32);
  OFF(sbr_alt_sin_twiddle_l32); OFF(sbr_t_cos_sin_l32); OFF(post_fft_tbl); OFF(dct23_tw); OFF(qmf_c);
  OFF(dig_rev_table2_128);
#undef OFF
  return n;
}

/* ixheaacd_cos_sin_mod (decoder/generic/ixheaacd_qmf_dec_generic.c:259): no_channels = 64 (synthesis, M = 32)
 * or 32 (analysis, M = 16); subband[128] in place. */
void ref_cos_sin_mod(int32_t *subband, int32_t no_channels) {
  ia_sbr_qmf_filter_bank_struct bank;
  ia_qmf_dec_tables_struct *t = (ia_qmf_dec_tables_struct *)&ixheaacd_aac_qmf_dec_tables;
  memset(&bank, 0, sizeof(bank));
  bank.no_channels = no_channels;
  if (no_channels == 64) {
    bank.cos_twiddle = t->sbr_sin_cos_twiddle_l64;
    bank.alt_sin_twiddle = t->sbr_alt_sin_twiddle_l64;
    ixheaacd_cos_sin_mod(subband, &bank, t->w_32, t->dig_rev_table2_32);
  } else {
    bank.cos_twiddle = t->sbr_sin_cos_twiddle_l32;
    bank.alt_sin_twiddle = t->sbr_alt_sin_twiddle_l32;
    ixheaacd_cos_sin_mod(subband, &bank, t->w_16, t->dig_rev_table4_16);
  }
}

/* ixheaacd_cplx_synt_qmffilt (decoder/ixheaacd_qmf_dec.c:811), HQ (complex) 64-band, no PS, no DRC, AAC-LC/HE-AAC
 * object type — exactly how ixheaacd_sbr_dec calls it at sbr_dec.c:1273 for a non-PS HQ channel.
 *   matrix        [32][128] : per slot re[64] | im[64]  (modified in place: block shifts + modulation)
 *   filter_states [1280] WORD16, in/out;  *drc_offset in/out (0..1152 step 128);  *filter_pos in/out (0..576 step 64)
 *   sf[4] = {ov_lb_scale, lb_scale, hb_scale, st_syn_scale}
 *   time_out      2048 WORD16 at stride ch_fac */
void ref_synt_qmffilt_hq(int32_t *matrix, int16_t *filter_states, int32_t *drc_offset, int32_t *filter_pos,
                         const int32_t *sf, int32_t lsb, int32_t usb, int32_t split, int16_t *time_out,
                         int32_t ch_fac) {
  ia_sbr_qmf_filter_bank_struct bank;
  ia_sbr_scale_fact_struct s;
  ia_sbr_tables_struct tabs;
  WORD32 dump[32][128]; /* qmf_real_out / qmf_imag_out sink (thread-private) */
  WORD32 *re[MAX_ENV_COLS], *im[MAX_ENV_COLS], *ore[MAX_ENV_COLS], *oim[MAX_ENV_COLS];
  ia_qmf_dec_tables_struct *t = (ia_qmf_dec_tables_struct *)&ixheaacd_aac_qmf_dec_tables;
  memset(&bank, 0, sizeof(bank));
  memset(&s, 0, sizeof(s));
  memset(&tabs, 0, sizeof(tabs));
  tabs.qmf_dec_tables_ptr = t;
  for (int i = 0; i < 32; i++) {
    re[i] = matrix + 128 * i;
    im[i] = re[i] + 64;
    ore[i] = dump[i];
    oim[i] = dump[i] + 64;
  }
  bank.no_channels = 64;
  bank.num_time_slots = 32;
  bank.lsb = (WORD16)lsb;
  bank.usb = (WORD16)usb;
  bank.filter_states = filter_states;
  bank.ixheaacd_drc_offset = (WORD16)*drc_offset;
  bank.p_filter = t->qmf_c;
  bank.filter_pos_syn = (WORD16 *)t->qmf_c + *filter_pos;
  s.ov_lb_scale = (WORD16)sf[0];
  s.lb_scale = (WORD16)sf[1];
  s.hb_scale = (WORD16)sf[2];
  s.st_syn_scale = (WORD16)sf[3];
  ixheaacd_cplx_synt_qmffilt(re, im, split, ore, oim, &s, time_out, &bank, NULL, 0, 0, &tabs, NULL, ch_fac, 0,
                             NULL, AOT_SBR);
  *drc_offset = bank.ixheaacd_drc_offset;
  *filter_pos = (int32_t)(bank.filter_pos_syn - (WORD16 *)t->qmf_c);
}

/* ixheaacd_cplx_anal_qmffilt (decoder/generic/ixheaacd_qmf_dec_generic.c:590), HQ 32-band, as called from
 * ixheaacd_sbr_dec (sbr_dec.c:1025).
 *   time_in   1024 WORD16 at stride ch_fac
 *   states    [320] WORD16 in/out; *pos in/out (offset of core_samples_buffer, 0..288 step 32);
 *   *filter_pos in/out (offset into qmf_c)
 *   matrix    [32][128]: real[i][0..31] at matrix[128 i], imag at +64 (only the first 32 of each written)
 *   returns lb_scale set by the stage */
int ref_anal_qmffilt_hq(const int16_t *time_in, int32_t ch_fac, int16_t *states, int32_t *pos, int32_t *filter_pos,
                        int32_t usb, int32_t *matrix) {
  ia_sbr_qmf_filter_bank_struct bank;
  ia_sbr_scale_fact_struct s;
  WORD32 *re[MAX_ENV_COLS], *im[MAX_ENV_COLS];
  ia_qmf_dec_tables_struct *t = (ia_qmf_dec_tables_struct *)&ixheaacd_aac_qmf_dec_tables;
  memset(&bank, 0, sizeof(bank));
  memset(&s, 0, sizeof(s));
  for (int i = 0; i < 32; i++) {
    re[i] = matrix + 128 * i;
    im[i] = re[i] + 64;
  }
  bank.no_channels = 32;
  bank.num_time_slots = 32;
  bank.lsb = 0;
  bank.usb = (WORD16)usb;
  bank.anal_filter_states = states;
  bank.core_samples_buffer = states + *pos;
  bank.analy_win_coeff = t->qmf_c;
  bank.filter_pos = (WORD16 *)t->qmf_c + *filter_pos;
  ixheaacd_cplx_anal_qmffilt(time_in, &s, re, im, &bank, t, ch_fac, 0, AOT_SBR);
  *pos = (int32_t)(bank.core_samples_buffer - states);
  *filter_pos = (int32_t)(bank.filter_pos - (WORD16 *)t->qmf_c);
  return s.lb_scale;
}

/* batch driver for the CPU baseline: n units, unit-major, private state per unit */
void ref_synt_qmffilt_hq_batch(int32_t *matrix, int16_t *filter_states, int32_t *drc_offset, int32_t *filter_pos,
                               const int32_t *sf, const int32_t *lsb, const int32_t *usb, int16_t *time_out,
                               int32_t n) {
  for (int32_t u = 0; u < n; u++)
    ref_synt_qmffilt_hq(matrix + (size_t)u * 4096, filter_states + (size_t)u * 1280, drc_offset + u, filter_pos + u,
                        sf + 4 * u, lsb[u], usb[u], 6, time_out + (size_t)u * 2048, 1);
}

/* ------------------------------------------------------------------------------------------------
 * HQ HF generator: ixheaacd_hf_generator (decoder/ixheaacd_lpp_tran.c:956-1258) as ixheaacd_sbr_dec calls it at
 * sbr_dec.c:1169.
 *   lpc      [2][128]  lpc_filt_states_{real,imag}[i] as re[64] | im[64] rows (read-only here)
 *   matrix   [38][128] QMF rows (6 overlap + 32 current), high band written in place
 *   prm      [80] WORD16, layout XO_HF_* in oracle/src/xaac_oracle.h
 *   bw_prev  [6] WORD32 bw_array_prev, in/out
 *   returns hb_scale
 * ---------------------------------------------------------------------------------------------- */
int ref_hf_generator_hq(const int32_t *lpc, int32_t *matrix, const int16_t *prm, int32_t *bw_prev) {
  ia_sbr_hf_generator_struct hf;
  ia_transposer_settings_struct set;
  ia_sbr_scale_fact_struct sf;
  WORD32 *re[40], *im[40];
  WORD32 invf[MAX_NUM_NOISE_VALUES], invf_prev[MAX_NUM_NOISE_VALUES];
  static __thread WORD32 scratch[40 * 128];
  static __thread WORD32 lpc_r[2][64], lpc_i[2][64];
  memset(&hf, 0, sizeof(hf));
  memset(&set, 0, sizeof(set));
  memset(&sf, 0, sizeof(sf));
  set.num_patches = prm[0];
  set.start_patch = prm[1];
  set.stop_patch = prm[2];
  set.num_columns = prm[3];
  for (int i = 0; i < MAX_NUM_NOISE_VALUES; i++) set.bw_borders[i] = prm[4 + i];
  for (int p = 0; p < MAX_NUM_PATCHES; p++) {
    const int16_t *q = prm + 14 + 6 * p;
    set.str_patch_param[p].src_start_band = q[0];
    set.str_patch_param[p].src_end_band = q[1];
    set.str_patch_param[p].guard_start_band = q[2];
    set.str_patch_param[p].dst_start_band = q[3];
    set.str_patch_param[p].dst_end_band = q[4];
    set.str_patch_param[p].num_bands_in_patch = q[5];
  }
  for (int i = 0; i < MAX_NUM_NOISE_VALUES; i++) {
    invf[i] = prm[54 + i];
    invf_prev[i] = prm[64 + i];
  }
  sf.ov_lb_scale = prm[74];
  sf.lb_scale = prm[75];
  hf.pstr_settings = &set;
  for (int i = 0; i < MAX_NUM_PATCHES; i++) hf.bw_array_prev[i] = bw_prev[i];
  for (int i = 0; i < 2; i++) {
    memcpy(lpc_r[i], lpc + 128 * i, 64 * sizeof(WORD32));
    memcpy(lpc_i[i], lpc + 128 * i + 64, 64 * sizeof(WORD32));
    hf.lpc_filt_states_real[i] = lpc_r[i];
    hf.lpc_filt_states_imag[i] = lpc_i[i];
  }
  memset(scratch, 0, 2 * 128 * sizeof(WORD32));
  memcpy(scratch + 256, matrix, 38 * 128 * sizeof(WORD32));
  for (int i = 0; i < 38; i++) {
    re[i] = scratch + 256 + 128 * i;
    im[i] = re[i] + 64;
  }
  /* factor (time_step), first_slot_offset, last_slot_offset are pre-multiplied by the caller of this shim:
   * prm[52] = start_idx argument (border_vec[0]), prm[53] = stop_idx argument (border_vec[num_env] - num_time_slots) */
  ixheaacd_hf_generator(&hf, &sf, re, im, prm[50], prm[52], prm[53], prm[51], prm[76], invf, invf_prev, scratch,
                        AOT_SBR);
  memcpy(matrix, scratch + 256, 38 * 128 * sizeof(WORD32));
  for (int i = 0; i < MAX_NUM_PATCHES; i++) bw_prev[i] = hf.bw_array_prev[i];
  return sf.hb_scale;
}

/* ------------------------------------------------------------------------------------------------
 * HQ envelope adjuster: ixheaacd_calc_sbrenvelope (decoder/ixheaacd_env_calc.c:692-1015) as ixheaacd_sbr_dec calls
 * it at sbr_dec.c:1195 with low_pow_flag = 0.  Record layouts: XO_ENV_* in oracle/src/xaac_oracle.h.
 *   prm [656] WORD16 side info; sf [8] WORD16 scale factors in/out; state [232] WORD16 in/out;
 *   matrix [38][128] in/out.  Returns the reference's error code.
 * ---------------------------------------------------------------------------------------------- */
#include "ref_pack.h"
const void *ref_rom_env_tables(int *bytes) {
  if (bytes) *bytes = (int)sizeof(ixheaacd_aac_dec_env_calc_tables);
  return &ixheaacd_aac_dec_env_calc_tables;
}
const void *ref_rom_misc_tables(int *bytes) {
  if (bytes) *bytes = (int)sizeof(ixheaacd_str_fft_n_transcendent_tables);
  return &ixheaacd_str_fft_n_transcendent_tables;
}
int ref_rom_env_offsets(int *o) {
  int n = 0;
  o[n++] = (int)offsetof(ia_env_calc_tables_struct, sbr_lim_gains_m);
  o[n++] = (int)offsetof(ia_env_calc_tables_struct, sbr_smooth_filter);
  o[n++] = (int)offsetof(ia_env_calc_tables_struct, sbr_inv_int_table);
  o[n++] = (int)offsetof(ia_env_calc_tables_struct, sbr_rand_ph);
  o[n++] = (int)sizeof(ia_env_calc_tables_struct);
  o[n++] = (int)offsetof(ixheaacd_misc_tables, inv_table);
  o[n++] = (int)offsetof(ixheaacd_misc_tables, sqrt_table);
  o[n++] = (int)offsetof(ixheaacd_misc_tables, dummy);
  return n;
}

int ref_calc_sbrenvelope_hq(const int16_t *prm, int16_t *sf, int16_t *state, int32_t *matrix) {
  static __thread struct { ia_sbr_frame_info_data_struct f; WORD16 extra[512]; } __attribute__((aligned(8))) fd;
  ia_sbr_header_data_struct h;
  ia_freq_band_data_struct fb;
  ia_sbr_prev_frame_data_struct pv;
  ia_sbr_scale_fact_struct s;
  ia_sbr_calc_env_struct ce;
  ia_sbr_tables_struct tabs;
  WORD16 filt_me[2 * MAX_FREQ_COEFFS], filt_noise[MAX_FREQ_COEFFS], deg[64];
  WORD32 *re[MAX_ENV_COLS], *im[MAX_ENV_COLS];
  memset(&fd, 0, sizeof(fd)); memset(&h, 0, sizeof(h)); memset(&fb, 0, sizeof(fb)); memset(&pv, 0, sizeof(pv));
  memset(&s, 0, sizeof(s)); memset(&ce, 0, sizeof(ce)); memset(&tabs, 0, sizeof(tabs)); memset(deg, 0, sizeof(deg));
  unpack_env_prm(prm, &h, &fb, &fd.f, &pv);
  unpack_sf(sf, &s);
  ce.filt_buf_me = filt_me;
  ce.filt_buf_noise_m = filt_noise;
  unpack_env_state(state, &ce);
  tabs.env_calc_tables_ptr = (ia_env_calc_tables_struct *)&ixheaacd_aac_dec_env_calc_tables;
  tabs.qmf_dec_tables_ptr = (ia_qmf_dec_tables_struct *)&ixheaacd_aac_qmf_dec_tables;
  tabs.sbr_rand_ph = tabs.env_calc_tables_ptr->sbr_rand_ph;
  for (int i = 0; i < 38; i++) { re[i] = matrix + 128 * i; im[i] = re[i] + 64; }
  IA_ERRORCODE err = ixheaacd_calc_sbrenvelope(&s, &ce, &h, &fd.f, &pv, re, im, deg, 0, &tabs,
                                               (ixheaacd_misc_tables *)&ixheaacd_str_fft_n_transcendent_tables, matrix,
                                               AOT_SBR);
  pack_sf(sf, &s);
  pack_env_state(state, &ce);
  return (int)err;
}

/* ------------------------------------------------------------------------------------------------
 * Whole fixed-point HQ SBR stage: ixheaacd_sbr_dec (decoder/ixheaacd_sbr_dec.c:662) with low_pow_flag = 0, driven
 * from the flat records of the C-ABI (XO_SIDE_*, XO_SBR_ST_*, XO_PS_ST_* in oracle/src/xaac_oracle.h).
 *   side [1232] WORD16; st [3920] WORD16 in/out; ps [3888] WORD16 in/out or NULL (mono SBR);
 *   time_in 1024 WORD16; out_l / out_r 2048 WORD16 each (out_r only with PS).  Returns the stage's return value.
 * ---------------------------------------------------------------------------------------------- */
const void *ref_rom_ps_tables(int *bytes) {
  if (bytes) *bytes = (int)sizeof(ixheaacd_aac_dec_ps_tables);
  return &ixheaacd_aac_dec_ps_tables;
}
int ref_rom_ps_offsets(int *o) {
  int n = 0;
#define OFF(m) o[n++] = (int)offsetof(ia_ps_tables_struct, m)
  OFF(decay_scale_factor); OFF(hyb_resol); OFF(rev_link_decay_ser); OFF(rev_link_delay_ser); OFF(borders_group);
  OFF(group_shift); OFF(group_to_bin); OFF(hybrid_to_bin); OFF(delay_to_bin); OFF(frac_delay_phase_fac_qmf_re_im);
  OFF(frac_delay_phase_fac_qmf_sub_re_im); OFF(frac_delay_phase_fac_qmf_ser_re_im);
  OFF(frac_delay_phase_fac_qmf_sub_ser_re_im); OFF(scale_factors); OFF(scale_factors_fine); OFF(alpha_values);
  OFF(p2_6); OFF(p8_13); OFF(huff_iid_dt);
#undef OFF
  return n;
}

int ref_sbr_dec_hq(const int16_t *side, int16_t *st, int16_t *ps, const int16_t *time_in, int16_t *out_l,
                   int16_t *out_r) {
  static __thread ref_sbr_ctx c;
  static __thread WORD16 tbuf[2 * 2048];
  const int use_ps = ps != NULL && side[XO_SIDE_PS];
  const int ch_fac = use_ps ? 2 : 1;
  unpack_sbr_ctx(&c, side, st, ps);
  memset(tbuf, 0, sizeof(tbuf));
  for (int i = 0; i < 1024; i++) tbuf[ch_fac * i] = time_in[i];
  WORD32 ret = ixheaacd_sbr_dec(&c.d, tbuf, &c.h, &c.fd.f, &c.pv, ps ? &c.ps : NULL, ps ? &c.bank_r : NULL,
                                ps ? &c.sf_r : NULL, side[XO_SIDE_APPLY], 0, c.work, &c.tabs,
                                (ixheaacd_misc_tables *)&ixheaacd_str_fft_n_transcendent_tables, ch_fac, NULL, 0, NULL,
                                AOT_SBR, 0, NULL, 0, 0);
  pack_sbr_state(st, &c.d, &c.pv);
  if (ps) pack_ps_state(ps, &c.ps, &c.bank_r, &c.sf_r);
  for (int i = 0; i < 2048; i++) out_l[i] = tbuf[ch_fac * i];
  if (out_r) for (int i = 0; i < 2048; i++) out_r[i] = use_ps ? tbuf[2 * i + 1] : 0;
  return (int)ret;
}

/* Per-slot PS work exactly as the left ixheaacd_cplx_synt_qmffilt call does it (decoder/ixheaacd_qmf_dec.c:1003-1030):
 * ixheaacd_init_rot_env at the PS envelope borders + ixheaacd_apply_ps (+ ixheaacd_shiftrountine) for the 32 slots of a
 * frame.  m [38][128] in/out (left), right [32][128] out; lb_scale / ps_scale enter through sf[8]. */
VOID ixheaacd_apply_ps(ia_ps_dec_struct *, WORD32 **, WORD32 **, WORD32 *, WORD32 *, ia_sbr_scale_fact_struct *, WORD16,
                       ia_sbr_tables_struct *, WORD);
VOID ixheaacd_init_rot_env(ia_ps_dec_struct *, WORD16, WORD16, ia_sbr_tables_struct *, const WORD16 *);
VOID ixheaacd_shiftrountine(WORD32 *, WORD32 *, WORD32, WORD32);
void ref_ps_apply_frame(const int16_t *side, const int16_t *st, int16_t *ps, const int16_t *sf, int32_t *m,
                        int32_t *right, int usb, int common_shift) {
  static __thread ref_sbr_ctx c;
  ia_sbr_scale_fact_struct s;
  WORD32 *re[40], *im[40];
  unpack_sbr_ctx(&c, side, st, ps);
  unpack_sf(sf, &s);
  for (int i = 0; i < 38; i++) { re[i] = m + 128 * i; im[i] = re[i] + 64; }
  int env = 0;
  for (int i = 0; i < 32; i++) {
    if (i == c.ps.border_position[env]) {
      ixheaacd_init_rot_env(&c.ps, (WORD16)env, (WORD16)usb, &c.tabs, ixheaacd_str_fft_n_transcendent_tables.trig_data);
      env++;
    }
    ixheaacd_apply_ps(&c.ps, &re[i], &im[i], right + 128 * i, right + 128 * i + 64, &s, (WORD16)i, &c.tabs, 32);
    if (common_shift) ixheaacd_shiftrountine(re[i], im[i], 64, common_shift);
  }
  pack_ps_state(ps, &c.ps, &c.bank_r, &c.sf_r);
}

/* ------------------------------------------------------------------------------------------------
 * CPU baseline for the HE-AAC chain (bench.py): n persistent decoder channels, each with its own reference structs,
 * stepped one frame at a time through the UNMODIFIED reference stages exactly as ixheaacd_dec_execute chains them:
 * ixheaacd_imdct_process -> the WORD32->WORD16 loop of ixheaacd_allocate_sbr_scr -> ixheaacd_sbr_dec.
 * Only the per-frame side info is refreshed between steps (the state lives in the reference's own structs).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  ref_sbr_ctx c;
  WORD32 overlap[512];
  int32_t prev_shape, prev_seq;
} ref_chain_unit;
typedef struct { int n; ref_chain_unit *u; } ref_chain;

void *ref_chain_create(int n, const int16_t *side, const int16_t *st, const int16_t *ps) {
  ref_chain *h = (ref_chain *)calloc(1, sizeof(ref_chain));
  h->n = n;
  h->u = (ref_chain_unit *)calloc((size_t)n, sizeof(ref_chain_unit));
  if (!h->u) { free(h); return NULL; }
  for (int i = 0; i < n; i++)
    unpack_sbr_ctx(&h->u[i].c, side + (size_t)i * XO_SIDE_WORDS, st + (size_t)i * XO_SBR_ST_WORDS,
                   ps ? ps + (size_t)i * XO_PS_ST_WORDS : NULL);
  return h;
}
void ref_chain_destroy(void *hv) {
  ref_chain *h = (ref_chain *)hv;
  if (h) { free(h->u); free(h); }
}
/* units [a, b): spec [n][1024] (destroyed), ics [n][2] {window_sequence, window_shape}, side [n][1232],
 * out [n][2048][2] (stereo) */
void ref_chain_step(void *hv, int a, int b, int32_t *spec, const uint8_t *ics, const int16_t *side, int16_t *out) {
  ref_chain *h = (ref_chain *)hv;
  WORD32 tbuf32[1024];
  WORD16 tbuf[2 * 2048];
  for (int i = a; i < b; i++) {
    ref_chain_unit *u = &h->u[i];
    const int16_t *sd = side + (size_t)i * XO_SIDE_WORDS;
    int adj = ref_imdct_process(spec + (size_t)i * 1024, u->overlap, &u->prev_shape, &u->prev_seq, ics[2 * i],
                                ics[2 * i + 1], tbuf32, 1);
    const int use_ps = sd[XO_SIDE_PS] != 0;
    const int ch_fac = use_ps ? 2 : 1;
    for (int k = 0; k < 1024; k++) tbuf[ch_fac * k] = ixheaac_round16(ixheaac_shl32_sat(tbuf32[k], adj));
    /* refresh the per-frame side info in the persistent structs */
    ia_freq_band_data_struct *fb = &u->c.fb;
    ia_sbr_prev_frame_data_struct pv_keep = u->c.pv;
    unpack_env_prm(sd + XO_SIDE_ENV, &u->c.h, fb, &u->c.fd.f, &u->c.pv);
    u->c.pv = pv_keep; /* previous-frame data is state, maintained by ixheaacd_sbr_dec itself */
    const int16_t *hf = sd + XO_SIDE_HF;
    unpack_hf_settings(hf, &u->c.set);
    fb->num_if_bands = hf[XO_HF_NUM_IF_BANDS];
    for (int k = 0; k < MAX_NUM_NOISE_VALUES; k++) u->c.fd.f.sbr_invf_mode[k] = hf[XO_HF_INVF + k];
    if (use_ps) unpack_side_ps(sd, &u->c.ps);
    ixheaacd_sbr_dec(&u->c.d, tbuf, &u->c.h, &u->c.fd.f, &u->c.pv, &u->c.ps, &u->c.bank_r, &u->c.sf_r, sd[XO_SIDE_APPLY], 0,
                     u->c.work, &u->c.tabs, (ixheaacd_misc_tables *)&ixheaacd_str_fft_n_transcendent_tables, ch_fac, NULL,
                     0, NULL, AOT_SBR, 0, NULL, 0, 0);
    int16_t *o = out + (size_t)i * 4096;
    if (use_ps) memcpy(o, tbuf, 4096 * sizeof(int16_t));
    else for (int k = 0; k < 2048; k++) { o[2 * k] = tbuf[k]; o[2 * k + 1] = tbuf[k]; }
  }
}
int ref_chain_unit_bytes(void) { return (int)sizeof(ref_chain_unit); }


/* ixheaacd_sbr_dec with low_pow_flag = 1 (real-valued SBR, what the reference runs for stereo HE-AACv1), driven from
 * the flat records.  st [3920] in/out (overlap: 6 real slots of 64 words; LPC rows: 32 real words), time_in 1024,
 * out 2048. */
int ref_sbr_dec_lp(const int16_t *side, int16_t *st, const int16_t *time_in, int16_t *out) {
  static __thread ref_sbr_ctx c;
  static __thread WORD16 tbuf[2048];
  unpack_sbr_ctx_lp(&c, side, st, NULL, 1);
  memset(tbuf, 0, sizeof(tbuf));
  memcpy(tbuf, time_in, 1024 * sizeof(WORD16));
  WORD32 ret = ixheaacd_sbr_dec(&c.d, tbuf, &c.h, &c.fd.f, &c.pv, NULL, NULL, NULL, side[XO_SIDE_APPLY], 1, c.work, &c.tabs,
                                (ixheaacd_misc_tables *)&ixheaacd_str_fft_n_transcendent_tables, 1, NULL, 0, NULL, AOT_SBR,
                                0, NULL, 0, 0);
  pack_sbr_state_lp(st, &c.d, &c.pv, 1);
  memcpy(out, tbuf, 2048 * sizeof(int16_t));
  return (int)ret;
}

/* ------------------------------------------------------------------------------------------------
 * CPU baseline for the stereo HE-AACv1 chain (bench.py, BASELINE configs[2]): n persistent decoder CHANNELS (units
 * 2k / 2k+1 = L / R of stream k) stepped through the UNMODIFIED reference stages ixheaacd_imdct_process -> WORD32->WORD16
 * hand-over -> ixheaacd_sbr_dec with low_pow_flag = 1.  out [n / 2][2048][2] interleaved like the reference's buffer.
 * ---------------------------------------------------------------------------------------------- */
void *ref_chain_lp_create(int n, const int16_t *side, const int16_t *st) {
  ref_chain *h = (ref_chain *)calloc(1, sizeof(ref_chain));
  h->n = n;
  h->u = (ref_chain_unit *)calloc((size_t)n, sizeof(ref_chain_unit));
  if (!h->u) { free(h); return NULL; }
  for (int i = 0; i < n; i++)
    unpack_sbr_ctx_lp(&h->u[i].c, side + (size_t)i * XO_SIDE_WORDS, st + (size_t)i * XO_SBR_ST_WORDS, NULL, 1);
  return h;
}
void ref_chain_lp_step(void *hv, int a, int b, int32_t *spec, const uint8_t *ics, const int16_t *side, int16_t *out) {
  ref_chain *h = (ref_chain *)hv;
  WORD32 tbuf32[1024];
  WORD16 tbuf[2048];
  for (int i = a; i < b; i++) {
    ref_chain_unit *u = &h->u[i];
    const int16_t *sd = side + (size_t)i * XO_SIDE_WORDS;
    int adj = ref_imdct_process(spec + (size_t)i * 1024, u->overlap, &u->prev_shape, &u->prev_seq, ics[2 * i],
                                ics[2 * i + 1], tbuf32, 1);
    for (int k = 0; k < 1024; k++) tbuf[k] = ixheaac_round16(ixheaac_shl32_sat(tbuf32[k], adj));
    ia_freq_band_data_struct *fb = &u->c.fb;
    ia_sbr_prev_frame_data_struct pv_keep = u->c.pv;
    unpack_env_prm(sd + XO_SIDE_ENV, &u->c.h, fb, &u->c.fd.f, &u->c.pv);
    u->c.pv = pv_keep;
    const int16_t *hf = sd + XO_SIDE_HF;
    unpack_hf_settings(hf, &u->c.set);
    fb->num_if_bands = hf[XO_HF_NUM_IF_BANDS];
    for (int k = 0; k < MAX_NUM_NOISE_VALUES; k++) u->c.fd.f.sbr_invf_mode[k] = hf[XO_HF_INVF + k];
    ixheaacd_sbr_dec(&u->c.d, tbuf, &u->c.h, &u->c.fd.f, &u->c.pv, NULL, NULL, NULL, sd[XO_SIDE_APPLY], 1, u->c.work,
                     &u->c.tabs, (ixheaacd_misc_tables *)&ixheaacd_str_fft_n_transcendent_tables, 1, NULL, 0, NULL, AOT_SBR,
                     0, NULL, 0, 0);
    int16_t *o = out + (size_t)(i >> 1) * 4096 + (i & 1);
    for (int k = 0; k < 2048; k++) o[2 * k] = tbuf[k];
  }
}

/* ------------------------------------------------------------------------------------------------
 * eSBR 64-band synthesis bank: the per-slot core of ixheaacd_esbr_synthesis_filt_block (decoder/ixheaacd_sbr_dec.c:
 * 583-654, stereo_config_idx <= 0, 64 channels) driven leaf by leaf in the reference's own order:
 * float -> WORD32 (x 64), ixheaacd_esbr_inv_modulation, ixheaacd_shiftrountine_with_rnd_hq, ixheaacd_esbr_qmfsyn64_winadd,
 * WORD32 -> float (/ 65536).  qmf [32][128] (re 64 | im 64), fs [1280] in/out, pos {drc_offset, filter_pos_syn_32 offset}.
 * ---------------------------------------------------------------------------------------------- */
VOID ixheaacd_esbr_inv_modulation(WORD32 *, ia_sbr_qmf_filter_bank_struct *, ia_qmf_dec_tables_struct *, WORD32);
VOID ixheaacd_shiftrountine_with_rnd_hq(WORD32 *, WORD32 *, WORD32 *, WORD32, WORD32);
VOID ixheaacd_esbr_qmfsyn64_winadd(WORD32 *, WORD32 *, WORD32 *, WORD32 *, WORD32);

const void *ref_rom_esbr_tables(int *bytes) {
  static int32_t blob[1280 + 60 + 64 + 32 + 24 + 32 + 16 + 64];
  const ia_qmf_dec_tables_struct *q = &ixheaacd_aac_qmf_dec_tables;
  memcpy(blob, q->esbr_qmf_c, 1280 * 4);
  memcpy(blob + 1280, q->esbr_w_32, 60 * 4);
  memcpy(blob + 1340, q->esbr_sin_cos_twiddle_l64, 64 * 4);
  memcpy(blob + 1404, q->esbr_alt_sin_twiddle_l64, 32 * 4);
  memcpy(blob + 1436, q->esbr_w_16, 24 * 4);
  memcpy(blob + 1460, q->esbr_sin_cos_twiddle_l32, 32 * 4);
  memcpy(blob + 1492, q->esbr_alt_sin_twiddle_l32, 16 * 4);
  memcpy(blob + 1508, q->esbr_t_cos_sin_l32, 64 * 4);
  if (bytes) *bytes = (int)sizeof(blob);
  return blob;
}

void ref_esbr_synth64(const float *qmf, int32_t *fs, int32_t *pos, float *out) {
  ia_qmf_dec_tables_struct *qt = (ia_qmf_dec_tables_struct *)&ixheaacd_aac_qmf_dec_tables;
  ia_sbr_qmf_filter_bank_struct bank;
  memset(&bank, 0, sizeof(bank));
  bank.no_channels = 64;
  bank.esbr_cos_twiddle = (WORD32 *)qt->esbr_sin_cos_twiddle_l64;
  bank.esbr_alt_sin_twiddle = (WORD32 *)qt->esbr_alt_sin_twiddle_l64;
  WORD32 *f1 = fs, *f2 = fs + 64, sixty4 = 64;
  WORD32 off = pos[0];
  WORD32 *filter_l = (WORD32 *)qt->esbr_qmf_c + pos[1];
  for (int i = 0; i < 32; i++) {
    WORD32 buf[128], time_out[64];
    for (int k = 0; k < 64; k++) {
      buf[k] = (WORD32)(qmf[128 * i + k] * 64);
      buf[k + 64] = (WORD32)(qmf[128 * i + 64 + k] * 64);
    }
    ixheaacd_esbr_inv_modulation(buf, &bank, qt, 64);
    ixheaacd_shiftrountine_with_rnd_hq(buf, buf + 64, &fs[off], 64, 6);
    ixheaacd_esbr_qmfsyn64_winadd(f1, f2, filter_l, time_out, 1);
    for (int k = 0; k < 64; k++) out[64 * i + k] = (FLOAT32)time_out[k] / (1 << 16);
    f1 += sixty4;
    f2 -= sixty4;
    sixty4 = -sixty4;
    off -= 128;
    if (off < 0) off += 1280;
    filter_l += 64;
    if (filter_l == (WORD32 *)qt->esbr_qmf_c + 640) filter_l = (WORD32 *)qt->esbr_qmf_c;
  }
  pos[0] = off;
  pos[1] = (int32_t)(filter_l - (WORD32 *)qt->esbr_qmf_c);
}
void ref_esbr_synth64_batch(const float *qmf, int32_t *fs, int32_t *pos, float *out, int n) {
  for (int u = 0; u < n; u++)
    ref_esbr_synth64(qmf + (size_t)u * 4096, fs + (size_t)u * 1280, pos + 2 * u, out + (size_t)u * 2048);
}


/* eSBR 32-band analysis bank: ixheaacd_esbr_analysis_filt_block (decoder/ixheaacd_sbr_dec.c:185-295) itself, on a minimal
 * ia_sbr_dec_struct: time_in [1024] float, states [320] in/out, pos {state_new_samples_pos_low_32 offset, filter_pos_32
 * offset} in/out, qmf [32][128] float out (re at +0..31, im at +64..95). */
VOID ixheaacd_esbr_analysis_filt_block(ia_sbr_dec_struct *, ia_sbr_tables_struct *, WORD32);
void ref_esbr_anal32(const float *time_in, int32_t *states, int32_t *pos, float *qmf) {
  static __thread ia_sbr_dec_struct d;
  ia_qmf_dec_tables_struct *qt = (ia_qmf_dec_tables_struct *)&ixheaacd_aac_qmf_dec_tables;
  ia_sbr_tables_struct tabs;
  memset(&tabs, 0, sizeof(tabs));
  memset(&d, 0, sizeof(d));
  tabs.qmf_dec_tables_ptr = qt;
  ia_sbr_qmf_filter_bank_struct *b = &d.str_codec_qmf_bank;
  b->no_channels = 32;
  b->num_time_slots = 32;
  b->lsb = 0;
  b->anal_filter_states_32 = states;
  b->state_new_samples_pos_low_32 = states + pos[0];
  b->analy_win_coeff_32 = qt->esbr_qmf_c;
  b->filter_pos_32 = (WORD32 *)qt->esbr_qmf_c + pos[1];
  b->esbr_cos_twiddle = (WORD32 *)qt->esbr_sin_cos_twiddle_l32;
  b->esbr_alt_sin_twiddle = (WORD32 *)qt->esbr_alt_sin_twiddle_l32;
  b->esbr_t_cos = (WORD32 *)qt->esbr_t_cos_sin_l32;
  d.time_sample_buf = (FLOAT32 *)time_in;
  ixheaacd_esbr_analysis_filt_block(&d, &tabs, 0);
  for (int i = 0; i < 32; i++)
    for (int k = 0; k < 32; k++) { qmf[128 * i + k] = d.qmf_buf_real[i][k]; qmf[128 * i + 64 + k] = d.qmf_buf_imag[i][k]; }
  pos[0] = (int32_t)(b->state_new_samples_pos_low_32 - states);
  pos[1] = (int32_t)(b->filter_pos_32 - (WORD32 *)qt->esbr_qmf_c);
}
void ref_esbr_anal32_batch(const float *time_in, int32_t *states, int32_t *pos, float *qmf, int n) {
  for (int u = 0; u < n; u++)
    ref_esbr_anal32(time_in + (size_t)u * 1024, states + (size_t)u * 320, pos + 2 * u, qmf + (size_t)u * 4096);
}

/* ---- eSBR float HF generator: the unmodified ixheaacd_generate_hf (decoder/ixheaacd_sbrdec_lpfuncs.c:981) on flat
 * records in the XO_EHF_* layout (oracle/src/xaac_oracle.h).  Buffers are [40][64]; the reference gets them + 2 rows. */
WORD32 ixheaacd_generate_hf(FLOAT32 ptr_src_buf_real[][64], FLOAT32 ptr_src_buf_imag[][64],
                            FLOAT32 ptr_ph_vocod_buf_real[][64], FLOAT32 ptr_ph_vocod_buf_imag[][64],
                            FLOAT32 ptr_dst_buf_real[][64], FLOAT32 ptr_dst_buf_imag[][64],
                            ia_sbr_frame_info_data_struct *ptr_frame_data, ia_sbr_header_data_struct *ptr_header_data,
                            WORD32 ldmps_present, WORD32 time_slots, WORD32 ec_flag);
int ref_esbr_generate_hf(const float *src_re, const float *src_im, const float *pv_re, const float *pv_im, float *dst_re,
                         float *dst_im, const int32_t *par, float *bw_prev, int32_t *patch_out) {
  static __thread ia_sbr_frame_info_data_struct fd;
  static __thread ia_sbr_header_data_struct hd;
  static __thread ia_freq_band_data_struct fb;
  memset(&fd, 0, sizeof(fd));
  memset(&hd, 0, sizeof(hd));
  memset(&fb, 0, sizeof(fb));
  hd.pstr_freq_band_data = &fb;
  fb.num_mf_bands = (WORD16)par[XO_EHF_NUM_MF];
  fb.num_nf_bands = (WORD16)par[XO_EHF_NUM_IF];
  fb.sub_band_start = (WORD16)par[XO_EHF_SB_START];
  for (int i = 0; i < 5; i++) fb.freq_band_tbl_noise[1 + i] = (WORD16)par[XO_EHF_INVF_TBL + i];
  for (int i = 0; i < 57; i++) fb.f_master_tbl[i] = (WORD16)par[XO_EHF_FMASTER + i];
  fd.str_frame_info_details.num_env = 1;
  fd.str_frame_info_details.border_vec[0] = (WORD16)par[XO_EHF_BORDER_FIRST];
  fd.str_frame_info_details.border_vec[1] = (WORD16)par[XO_EHF_BORDER_LAST];
  hd.hbe_flag = par[XO_EHF_HBE_FLAG];
  fd.sbr_patching_mode = par[XO_EHF_PATCHING_MODE];
  hd.out_sampling_freq = par[XO_EHF_FS];
  hd.pre_proc_flag = par[XO_EHF_PRE_PROC];
  hd.is_usf_4 = par[XO_EHF_USF4];
  fd.mps_sbr_flag = par[XO_EHF_MPS_SBR];
  fd.cov_count = par[XO_EHF_COV_COUNT];
  for (int i = 0; i < 5; i++) {
    fd.sbr_invf_mode[i] = par[XO_EHF_INVF + i];
    fd.sbr_invf_mode_prev[i] = par[XO_EHF_INVF_PREV + i];
  }
  for (int i = 0; i < 6; i++) fd.bw_array_prev[i] = bw_prev[i];
  typedef FLOAT32(*rows_t)[64];
  WORD32 e = ixheaacd_generate_hf((rows_t)(src_re + 128), (rows_t)(src_im + 128), pv_re ? (rows_t)(pv_re + 128) : NULL,
                                  pv_im ? (rows_t)(pv_im + 128) : NULL, (rows_t)(dst_re + 128), (rows_t)(dst_im + 128),
                                  &fd, &hd, 0, 32, 0);
  if (e) return e;
  patch_out[0] = fd.patch_param.num_patches;
  for (int i = 0; i < 7; i++) patch_out[1 + i] = fd.patch_param.start_subband[i];
  for (int i = 0; i < 6; i++) bw_prev[i] = fd.bw_array_prev[i];
  return 0;
}
void ref_esbr_generate_hf_batch(const float *src_re, const float *src_im, const float *pv_re, const float *pv_im,
                                float *dst_re, float *dst_im, const int32_t *par, float *bw_prev, int32_t *patch_out,
                                int32_t *err, int n) {
  const size_t B = (size_t)XO_EHF_ROWS * 64;
  for (int u = 0; u < n; u++)
    err[u] = ref_esbr_generate_hf(src_re + u * B, src_im + u * B, pv_re ? pv_re + u * B : 0, pv_im ? pv_im + u * B : 0,
                                  dst_re + u * B, dst_im + u * B, par + (size_t)u * XO_EHF_PAR_WORDS, bw_prev + 6 * u,
                                  patch_out + 8 * u);
}

/* ---- eSBR float envelope adjuster: the unmodified ixheaacd_sbr_env_calc (decoder/ixheaacd_esbr_envcal.c:71), ORIG_SBR
 * branch, on flat records in the XO_EEC_* layout.  re / im are [40][64]; the reference gets them + 2 rows. */
WORD32 ixheaacd_sbr_env_calc(ia_sbr_frame_info_data_struct *frame_data, FLOAT32 input_real[][64], FLOAT32 input_imag[][64],
                             FLOAT32 input_real1[][64], FLOAT32 input_imag1[][64], WORD32 x_over_qmf[MAX_NUM_PATCHES],
                             FLOAT32 *scratch_buff, FLOAT32 *env_out, WORD32 ldmps_present, WORD32 ec_flag);
extern const FLOAT32 ixheaac_random_phase[512][2];
void ref_rom_esbr_random_phase(float *out) { memcpy(out, ixheaac_random_phase, 1024 * sizeof(float)); }
static __thread const float *g_ec_low_re, *g_ec_low_im; /* set by ref_esbr_env_calc_tes: the low band inter-TES reads */
int ref_esbr_env_calc(float *re, float *im, int32_t *ipar, const float *fpar, float *state) {
  static __thread ia_sbr_frame_info_data_struct fd;
  static __thread ia_sbr_header_data_struct hd;
  static __thread ia_freq_band_data_struct fb;
  static __thread FLOAT32 scratch[1024], low_re[40][64], low_im[40][64];
  WORD32 x_over[MAX_NUM_PATCHES] = {0};
  memset(&fd, 0, sizeof(fd));
  memset(&hd, 0, sizeof(hd));
  memset(&fb, 0, sizeof(fb));
  fd.pstr_sbr_header = &hd;
  hd.pstr_freq_band_data = &fb;
  fb.sub_band_start = (WORD16)ipar[XO_EEC_SB_START];
  fb.sub_band_end = (WORD16)ipar[XO_EEC_SB_END];
  fb.num_sf_bands[0] = (WORD16)ipar[XO_EEC_NUM_SF_LO];
  fb.num_sf_bands[1] = (WORD16)ipar[XO_EEC_NUM_SF_HI];
  fb.num_nf_bands = (WORD16)ipar[XO_EEC_NUM_NF];
  fb.freq_band_table[0] = fb.freq_band_tbl_lo;
  fb.freq_band_table[1] = fb.freq_band_tbl_hi;
  for (int i = 0; i < 6; i++) fb.freq_band_tbl_noise[i] = (WORD16)ipar[XO_EEC_TBL_NOISE + i];
  for (int i = 0; i < 29; i++) fb.freq_band_tbl_lo[i] = (WORD16)ipar[XO_EEC_TBL_LO + i];
  for (int i = 0; i < 57; i++) fb.freq_band_tbl_hi[i] = (WORD16)ipar[XO_EEC_TBL_HI + i];
  hd.is_usf_4 = ipar[XO_EEC_USF4];
  hd.smoothing_mode = (WORD16)ipar[XO_EEC_SMOOTHING_MODE];
  hd.interpol_freq = (WORD16)ipar[XO_EEC_INTERPOL_FREQ];
  hd.limiter_bands = (WORD16)ipar[XO_EEC_LIMITER_BANDS];
  hd.limiter_gains = (WORD16)ipar[XO_EEC_LIMITER_GAINS];
  hd.esbr_start_up = ipar[XO_EEC_START_UP];
  hd.esbr_start_up_pvc = ipar[XO_EEC_START_UP];
  ia_frame_info_struct *fi = &fd.str_frame_info_details;
  fi->num_env = (WORD16)ipar[XO_EEC_NUM_ENV];
  fi->transient_env = (WORD16)ipar[XO_EEC_TRANS_ENV];
  fi->num_noise_env = (WORD16)ipar[XO_EEC_NUM_NOISE_ENV];
  for (int i = 0; i < 9; i++) fi->border_vec[i] = (WORD16)ipar[XO_EEC_BORDER + i];
  for (int i = 0; i < 8; i++) fi->freq_res[i] = (WORD16)ipar[XO_EEC_FREQ_RES + i];
  for (int i = 0; i < 3; i++) fi->noise_border_vec[i] = (WORD16)ipar[XO_EEC_NOISE_BORDER + i];
  for (int i = 0; i < 8; i++) fd.inter_temp_shape_mode[i] = ipar[XO_EEC_INTER_TES + i];
  for (int i = 0; i < 4; i++) fd.gate_mode[i] = ipar[XO_EEC_GATE_MODE + i];
  for (int i = 0; i < 52; i++) fd.lim_table[i / 13][i % 13] = ipar[XO_EEC_LIM_TABLE + i];
  for (int i = 0; i < 56; i++) fd.add_harmonics[i] = ipar[XO_EEC_ADD_HARM + i];
  memcpy(fd.harm_flag_prev, ipar + XO_EEC_HARM_PREV, 64);
  fd.env_short_flag_prev = ipar[XO_EEC_SHORT_PREV];
  fd.harm_index = ipar[XO_EEC_HARM_INDEX];
  fd.phase_index = ipar[XO_EEC_PHASE_INDEX];
  fd.reset_flag = ipar[XO_EEC_RESET];
  fd.sbr_mode = ipar[XO_EEC_SBR_MODE];
  fd.prev_sbr_mode = ipar[XO_EEC_SBR_MODE];
  fd.sbr_patching_mode = fd.prev_sbr_patching_mode = 0;
  memcpy(fd.flt_env_sf_arr, fpar + XO_EEC_SFB_NRG, 448 * sizeof(float));
  memcpy(fd.flt_noise_floor, fpar + XO_EEC_NOISE_FLOOR, 10 * sizeof(float));
  memcpy(fd.e_gain, state, 320 * sizeof(float));
  memcpy(fd.noise_buf, state + 320, 320 * sizeof(float));
  typedef FLOAT32(*rows_t)[64];
  if (g_ec_low_re) {
    memcpy(low_re, g_ec_low_re, sizeof(low_re));
    memcpy(low_im, g_ec_low_im, sizeof(low_im));
  }
  WORD32 e = ixheaacd_sbr_env_calc(&fd, (rows_t)(re + 128), (rows_t)(im + 128), low_re + 2, low_im + 2, x_over, scratch, NULL, 0, 0);
  if (e) return e;
  memcpy(ipar + XO_EEC_HARM_PREV, fd.harm_flag_prev, 64);
  ipar[XO_EEC_SHORT_PREV] = fd.env_short_flag_prev;
  ipar[XO_EEC_HARM_INDEX] = fd.harm_index;
  ipar[XO_EEC_PHASE_INDEX] = fd.phase_index;
  ipar[XO_EEC_START_UP] = hd.esbr_start_up;
  memcpy(state, fd.e_gain, 320 * sizeof(float));
  memcpy(state + 320, fd.noise_buf, 320 * sizeof(float));
  return 0;
}
/* with the low band (qmf_buf_real / imag rows 0..39) for envelopes that use inter-TES */
void ref_esbr_env_calc_tes_batch(float *re, float *im, const float *low_re, const float *low_im, int32_t *ipar, const float *fpar,
                                 float *state, int32_t *err, int n) {
  for (int u = 0; u < n; u++) {
    g_ec_low_re = low_re + (size_t)u * 2560;
    g_ec_low_im = low_im + (size_t)u * 2560;
    err[u] = ref_esbr_env_calc(re + (size_t)u * 2560, im + (size_t)u * 2560, ipar + (size_t)u * XO_EEC_IPAR_WORDS,
                               fpar + (size_t)u * XO_EEC_FPAR_WORDS, state + (size_t)u * XO_EEC_STATE_WORDS);
  }
  g_ec_low_re = g_ec_low_im = NULL;
}
void ref_esbr_env_calc_batch(float *re, float *im, int32_t *ipar, const float *fpar, float *state, int32_t *err, int n) {
  for (int u = 0; u < n; u++)
    err[u] = ref_esbr_env_calc(re + (size_t)u * 2560, im + (size_t)u * 2560, ipar + (size_t)u * XO_EEC_IPAR_WORDS,
                               fpar + (size_t)u * XO_EEC_FPAR_WORDS, state + (size_t)u * XO_EEC_STATE_WORDS);
}


/* ---- ixheaacd_samples_sat (decoder/ixheaacd_decode_main.c:82): float[ch][4096] -> interleaved PCM16 */
VOID ixheaacd_samples_sat(WORD8 *outbuffer, WORD32 num_samples_out, WORD32 pcmsize, FLOAT32 (*out_samples)[4096],
                          WORD32 *out_bytes, WORD32 num_channel_out);
void ref_samples_sat16(const float *in /* [nch][n] */, int nch, int n, int16_t *pcm) {
  static __thread FLOAT32 buf[8][4096];
  WORD32 bytes = 0;
  for (int c = 0; c < nch; c++) memcpy(buf[c], in + (size_t)c * n, n * sizeof(float));
  ixheaacd_samples_sat((WORD8 *)pcm, n, 16, buf, &bytes, nch);
}

/* ---- xHE-AAC chain of one channel unit on the reference's own functions (bench.py --workload xheaac_stereo_chain, CPU arm):
 * ixheaacd_fd_frm_dec -> x 2^-15 (ext_ch_ele.c:1040) -> the eSBR branch of ixheaacd_sbr_dec spelled out with its leaf
 * functions (history memmoves, ixheaacd_esbr_analysis_filt_block, ixheaacd_generate_hf, ixheaacd_sbr_env_calc,
 * regrouping as ixheaacd_esbr_synthesis_regrp, synthesis core) -> ixheaacd_samples_sat.  The analysis output is copied into
 * the stage arrays (the reference writes them in place: 8 KB of extra copy per unit in this arm). */
int ref_usac_fd_frm_dec(int32_t *coef, int32_t *overlap, int win_seq, int win_shape, int win_shape_prev, int32_t *out);
void ref_xheaac_chain_batch(int32_t *coef, int32_t *overlap, const int32_t *win_seq, const int32_t *win_shape,
                            const int32_t *win_shape_prev, float *q4 /* [n][4][2560] */, int32_t *anal, int32_t *apos,
                            int32_t *synth, int32_t *spos, float *bw, int32_t *patch, float *ec, const int32_t *hf_par,
                            int32_t *ec_ipar, const float *ec_fpar, const int32_t *rg, int16_t *pcm /* [n/2][2048][2] */,
                            int32_t *err, int a, int b) {
  static __thread int32_t core[1024];
  static __thread float tin[1024], qa[32 * 128], m[32 * 128], tout[2048];
  for (int u = a; u < b; u++) {
    float *q = q4 + (size_t)u * 4 * 2560;
    int e = ref_usac_fd_frm_dec(coef + (size_t)u * 1024, overlap + (size_t)u * 1024, win_seq[u], win_shape[u],
                                win_shape_prev[u], core);
    for (int k = 0; k < 1024; k++) tin[k] = (FLOAT32)((FLOAT32)core[k] * (FLOAT32)(0.000030517578125));
    for (int z = 0; z < 4; z++) memmove(q + 2560 * z, q + 2560 * z + 32 * 64, 8 * 64 * sizeof(float));
    ref_esbr_anal32(tin, anal + (size_t)u * 320, apos + 2 * u, qa);
    for (int s = 0; s < 32; s++) {
      memcpy(q + 64 * (8 + s), qa + 128 * s, 32 * sizeof(float));
      memcpy(q + 2560 + 64 * (8 + s), qa + 128 * s + 64, 32 * sizeof(float));
    }
    e |= ref_esbr_generate_hf(q, q + 2560, NULL, NULL, q + 5120, q + 7680, hf_par + (size_t)u * XO_EHF_PAR_WORDS, bw + 6 * u,
                              patch + 8 * u);
    e |= ref_esbr_env_calc(q + 5120, q + 7680, ec_ipar + (size_t)u * XO_EEC_IPAR_WORDS, ec_fpar + (size_t)u * XO_EEC_FPAR_WORDS,
                           ec + (size_t)u * 640);
    const int32_t *r = rg + 4 * u;
    for (int s = 0; s < 32; s++) {
      const int xo = s < r[2] ? r[0] : r[1];
      for (int k = 0; k < 64; k++) {
        m[128 * s + k] = k < xo ? q[64 * (2 + s) + k] : q[5120 + 64 * (2 + s) + k];
        m[128 * s + 64 + k] = k < xo ? q[2560 + 64 * (2 + s) + k] : q[7680 + 64 * (2 + s) + k];
      }
    }
    ref_esbr_synth64(m, synth + (size_t)u * 1280, spos + 2 * u, tout);
    int16_t *o = pcm + (size_t)(u / 2) * 4096 + (u & 1);
    for (int i = 0; i < 2048; i++) {
      float v = tout[i];
      if (v > 32767.0f) v = 32767.0f; else if (v < -32768.0f) v = -32768.0f;
      o[2 * i] = (int16_t)v;
    }
    err[u] = e;
  }
}
