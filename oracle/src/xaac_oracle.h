/*
 * oracle/src/xaac_oracle.h — TEST INFRASTRUCTURE ONLY.
 * Public surface of the CPU oracle (liboracle.so). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load it. The product library never links or calls it.
 */
#ifndef XAAC_ORACLE_H
#define XAAC_ORACLE_H
#include <stdint.h>

/* window_sequence codes — decoder/ixheaacd_cnst.h:100-103 */
enum { XO_ONLY_LONG = 0, XO_LONG_START = 1, XO_EIGHT_SHORT = 2, XO_LONG_STOP = 3 };

/* Byte offsets inside the IMDCT ROM blob = the leading 7500 bytes of ia_aac_dec_imdct_tables_struct
 * (decoder/ixheaacd_aac_rom.h:112-121). Same layout the product takes in xaac_b200_set_imdct_rom(). */
#define XO_ROM_COS 0             /* WORD16[514] */
#define XO_ROM_DIGREV_LONG 1028  /* WORD8[64]   */
#define XO_ROM_DIGREV_SHORT 1092 /* WORD8[8]    */
#define XO_ROM_FFT_TW 1100       /* WORD32[448] */
#define XO_ROM_WIN_LONG_SINE 2892
#define XO_ROM_WIN_LONG_KBD 4940
#define XO_ROM_WIN_SHORT_SINE 6988
#define XO_ROM_WIN_SHORT_KBD 7244
#define XO_ROM_IMDCT_BYTES 7500

int xo_calc_max_spectral_line(const int32_t *x, int n);
int xo_inverse_transform(const uint8_t *rom, int32_t *spec, int32_t *scratch, int expo, int n);
void xo_post_twiddle(const uint8_t *rom, int32_t *out, const int32_t *y, int n);
int xo_imdct_process(const uint8_t *rom, int32_t *spec, int32_t *ovl, int32_t *prev_shape, int32_t *prev_seq,
                     int win_seq, int win_shape, int32_t *out, int ch_fac);
void xo_imdct_process_batch(const uint8_t *rom, int32_t *spec, int32_t *ovl, int32_t *prev_shape,
                            int32_t *prev_seq, const int32_t *win_seq, const int32_t *win_shape, int32_t *out,
                            int32_t *qshift_adj, int n);
#endif
