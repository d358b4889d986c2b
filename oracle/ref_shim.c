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
