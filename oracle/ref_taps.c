/*
 * oracle/ref_taps.c — TEST INFRASTRUCTURE ONLY.
 *
 * Stage taps for the UNMODIFIED reference decoder, installed at link time with `ld --wrap`:
 * the reference testbench (test/decoder/*.c) + libxaacdec.a + this file give oracle/_ref/xaacdec_tap, which
 * decodes a real bitstream exactly like xaacdec and, when XAAC_TAP_FILE is set, appends one binary record
 * per stage call (inputs, state-before, outputs, state-after).  tools/make_golden.py turns those records
 * into the small fixtures under tests/golden/.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#define REF_SHIM_HEADERS_ONLY
#include "ref_headers.h"

static FILE *tap_fp(void) {
  static FILE *fp = NULL;
  static int tried = 0;
  if (!tried) {
    const char *p = getenv("XAAC_TAP_FILE");
    tried = 1;
    if (p && *p) fp = fopen(p, "wb");
  }
  return fp;
}
static int tap_limit(void) {
  static int lim = -1;
  if (lim < 0) {
    const char *p = getenv("XAAC_TAP_MAX");
    lim = p ? atoi(p) : 1000000;
  }
  return lim;
}

/* ---- ixheaacd_imdct_process (decoder/ixheaacd_lpfuncs.c:347) -------------------------------------------
 * record: int32 magic 'IMD1', int32 hdr[8] = {frame_length, object_type, ch_fac, prev_shape, prev_seq,
 *         win_seq, win_shape, qshift_adj}, int32 spec[1024], ovl_before[512], out[1024], ovl_after[512] */
VOID __real_ixheaacd_imdct_process(ia_aac_dec_overlap_info *, WORD32 *, ia_ics_info_struct *, VOID *,
                                   const WORD16, WORD32 *, ia_aac_dec_tables_struct *, WORD32, WORD32, WORD);

VOID __wrap_ixheaacd_imdct_process(ia_aac_dec_overlap_info *ovl, WORD32 *spec, ia_ics_info_struct *ics,
                                   VOID *out, const WORD16 ch_fac, WORD32 *scratch,
                                   ia_aac_dec_tables_struct *tabs, WORD32 object_type, WORD32 ld_mps,
                                   WORD slot) {
  static int count = 0;
  FILE *fp = tap_fp();
  int rec = fp && ics->frame_length == 1024 && count < tap_limit();
  int32_t hdr[9];
  int32_t spec_in[1024], ovl_in[512];
  if (rec) {
    hdr[0] = 0x31444d49;
    hdr[1] = ics->frame_length;
    hdr[2] = object_type;
    hdr[3] = ch_fac;
    hdr[4] = ovl->window_shape;
    hdr[5] = ovl->window_sequence;
    hdr[6] = ics->window_sequence;
    hdr[7] = ics->window_shape;
    memcpy(spec_in, spec, sizeof(spec_in));
    memcpy(ovl_in, ovl->ptr_overlap_buf, sizeof(ovl_in));
  }
  __real_ixheaacd_imdct_process(ovl, spec, ics, out, ch_fac, scratch, tabs, object_type, ld_mps, slot);
  if (rec) {
    int32_t o[1024];
    const WORD32 *po = (const WORD32 *)out;
    for (int i = 0; i < 1024; i++) o[i] = po[ch_fac * i];
    hdr[8] = ics->qshift_adj;
    fwrite(hdr, 4, 9, fp);
    fwrite(spec_in, 4, 1024, fp);
    fwrite(ovl_in, 4, 512, fp);
    fwrite(o, 4, 1024, fp);
    fwrite(ovl->ptr_overlap_buf, 4, 512, fp);
    fflush(fp);
    count++;
  }
}
