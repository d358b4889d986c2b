/*
 * oracle/src/fixmath.h — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain-C restatement of the fixed-point primitive semantics the libxaac hot path is composed of.
 * Each primitive cites the reference definition it follows (paths relative to /root/reference).
 * Compile with -fwrapv: signed overflow wraps, >> on negatives is arithmetic (reference: cmake/utils.cmake:17).
 */
#ifndef XAAC_ORACLE_FIXMATH_H
#define XAAC_ORACLE_FIXMATH_H
#include <stdint.h>

typedef int16_t i16;
typedef int32_t i32;
typedef int64_t i64;
typedef uint32_t u32;

#define OX_MAX32 ((i32)0x7fffffff)
#define OX_MIN32 ((i32)0x80000000)

static inline i32 ox_sat64(i64 v) { return v > OX_MAX32 ? OX_MAX32 : (v < OX_MIN32 ? OX_MIN32 : (i32)v); }

/* common/ixheaac_basic_ops40.h:34  (32x16 -> high 32 of the 48-bit product, floor) */
static inline i32 ox_mul32x16(i32 a, i16 b) { return (i32)(((i64)a * (i64)b) >> 16); }
/* decoder/ixheaacd_aac_imdct.c:80  (multiplier = low half of a packed word) */
static inline i32 ox_mul32x16l(i32 a, i32 b) { return (i32)(((i64)a * (i64)(i16)b) >> 16); }
/* common/ixheaac_basic_ops32.h:134 (multiplier = high half of a packed word) */
static inline i32 ox_mul32x16h(i32 a, i32 b) { return (i32)(((i64)a * (i64)(b >> 16)) >> 16); }
/* common/ixheaac_basic_ops32.h:144 and decoder/ixheaacd_aac_imdct.c:95 — FULL product, saturated, no shift */
static inline i32 ox_mul32x16_fullsat(i32 a, i16 b) { return ox_sat64((i64)a * (i64)b); }
/* common/ixheaac_basic_ops40.h:23 */
static inline i32 ox_mul32x16_shl(i32 a, i16 b) { return (i32)((u32)ox_mul32x16(a, b) << 1); }
/* common/ixheaac_basic_ops40.h:78 / :68 */
static inline i32 ox_mul32(i32 a, i32 b) { return (i32)(((i64)a * (i64)b) >> 32); }
static inline i32 ox_mul32_shl(i32 a, i32 b) { return (i32)((u32)ox_mul32(a, b) << 1); }

/* common/ixheaac_basic_ops32.h:181,189 — wrapping */
static inline i32 ox_add(i32 a, i32 b) { return (i32)((u32)a + (u32)b); }
static inline i32 ox_sub(i32 a, i32 b) { return (i32)((u32)a - (u32)b); }
static inline i32 ox_shl1(i32 a) { return (i32)((u32)a << 1); }
static inline i32 ox_lsl(i32 a, int s) { return (i32)((u32)a << s); }
/* common/ixheaac_basic_ops32.h:197,225 */
static inline i32 ox_add_sat(i32 a, i32 b) { return ox_sat64((i64)a + (i64)b); }
static inline i32 ox_sub_sat(i32 a, i32 b) { return ox_sat64((i64)a - (i64)b); }
/* common/ixheaac_basic_ops32.h:317 */
static inline i32 ox_neg_sat(i32 a) { return a == OX_MIN32 ? OX_MAX32 : -a; }
/* common/ixheaac_basic_ops32.h:283 */
static inline i32 ox_abs_nrm(i32 a) { return a < 0 ? ~a : a; }
/* common/ixheaac_basic_ops32.h:236 — redundant sign bits; 31 for 0 and -1 */
static inline int ox_norm32(i32 a) {
  if (a == 0 || a == -1) return 31;
  if (a < 0) a = ~a;
  int n = 0;
  while (a < (i32)0x40000000) { a <<= 1; n++; }
  return n;
}
/* common/ixheaac_basic_ops32.h:257 — same without the sign fold (callers pass non-negative) */
static inline int ox_pnorm32(i32 a) {
  if (a == 0) return 31;
  int n = 0;
  while (a < (i32)0x40000000) { a = (i32)((u32)a << 1); n++; if (n > 64) break; }
  return n;
}
/* common/ixheaac_basic_ops32.h:39 — shift count masked to 8 bits, >31 gives 0, wraps */
static inline i32 ox_shl32(i32 a, int b) { b &= 0xff; return b > 31 ? 0 : (i32)((u32)a << b); }
/* common/ixheaac_basic_ops32.h:51 */
static inline i32 ox_shr32(i32 a, int b) { b &= 0xff; return b >= 31 ? (a < 0 ? -1 : 0) : (a >> b); }
/* common/ixheaac_basic_ops32.h:67 — callers guarantee 0 <= b <= 31 */
static inline i32 ox_shl32_sat(i32 a, int b) {
  if (a > (OX_MAX32 >> b)) return OX_MAX32;
  if (a < (OX_MIN32 >> b)) return OX_MIN32;
  return (i32)((u32)a << b);
}
/* common/ixheaac_basic_ops32.h:377 — ROUNDING right shift */
static inline i32 ox_shr32_sat(i32 a, int b) {
  b &= 0xff;
  if (b >= 31) return a < 0 ? -1 : 0;
  if (b <= 0) return a;
  return ox_add_sat(a, (i32)1 << (b - 1)) >> b;
}
/* common/ixheaac_basic_ops.h:114 */
static inline i32 ox_shl32_dir_sat_limit(i32 a, int b) {
  if (b < 0) { b = -b; if (b > 31) b = 31; return ox_shr32(a, b); }
  return ox_shl32_sat(a, b);
}
/* common/ixheaac_basic_ops32.h:59-92 (shl32_dir / shr32_dir / *_dir_sat) */
static inline i32 ox_shl32_dir(i32 a, int b) { return b < 0 ? ox_shr32(a, -b) : ox_shl32(a, b); }
static inline i32 ox_shr32_dir(i32 a, int b) { return b < 0 ? ox_shl32(a, -b) : ox_shr32(a, b); }
static inline i32 ox_shl32_dir_sat(i32 a, int b) { return b < 0 ? ox_shr32(a, -b) : ox_shl32_sat(a, b); }
static inline i32 ox_shr32_dir_sat(i32 a, int b) { return b < 0 ? ox_shl32_sat(a, -b) : ox_shr32(a, b); }
/* common/ixheaac_basic_ops.h:104 */
static inline i32 ox_shr32_dir_sat_limit(i32 a, int b) {
  if (b < 0) return ox_shl32_sat(a, -b);
  if (b > 31) b = 31;
  return ox_shr32(a, b);
}

/* 16-bit family: common/ixheaac_basic_ops16.h */
static inline i16 ox_sat16(i32 v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : (i16)v); } /* :23 */
static inline i16 ox_neg16(i16 a) { return a == -32768 ? 32767 : (i16)(-a); }                      /* :205 */
static inline i16 ox_round16(i32 a) { return (i16)(ox_add_sat(a, 0x8000) >> 16); }                 /* :231 */
static inline i16 ox_add16(i16 a, i16 b) { return (i16)(a + b); }                                  /* :36 */
static inline i16 ox_sub16(i16 a, i16 b) { return (i16)(a - b); }                                  /* :48 */
static inline i16 ox_add16_sat(i16 a, i16 b) { return ox_sat16((i32)a + b); }                      /* :41 */
static inline i16 ox_sub16_sat(i16 a, i16 b) { return ox_sat16((i32)a - b); }                      /* :53 */
static inline i16 ox_mult16(i16 a, i16 b) { return (i16)(((i32)a * b) >> 16); }                    /* :59 */
static inline i16 ox_mult16_shl(i16 a, i16 b) { return (i16)(((i32)a * b) >> 15); }                /* :64 */
static inline i16 ox_mult16_shl_sat(i16 a, i16 b) { return ox_sat16(((i32)a * b) >> 15); }         /* :69 */
static inline i16 ox_shl16(i16 a, int s) { return (i16)(a << s); }                                 /* :76 */
static inline i16 ox_shr16(i16 a, int s) { return (i16)(a >> s); }                                 /* :91 */
static inline i16 ox_shl16_sat(i16 a, int s) { if (s > 15) s = 15; return ox_sat16((i32)(a << s)); } /* :81 */
static inline int ox_norm16(i16 a) {                                                               /* :136 */
  if (a == 0) return 0;
  if (a == -1) return 15;
  if (a < 0) a = (i16)~a;
  int n = 0;
  while (a < 0x4000) { a <<= 1; n++; }
  return n;
}
/* common/ixheaac_basic_ops32.h:125 etc. */
static inline i32 ox_mult16x16(i16 a, i16 b) { return (i32)a * (i32)b; }

#endif
