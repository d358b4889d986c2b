// fixmath.cuh — device-side fixed-point primitives for the libxaac hot path (sm_100a).
//
// Bit-exact equivalents of the reference's L1 arithmetic (paths relative to /root/reference):
//   common/ixheaac_basic_ops32.h, ixheaac_basic_ops40.h, ixheaac_basic_ops16.h, ixheaac_basic_ops.h.
// The reference is built with -fwrapv, so plain adds/subs/shifts wrap; we do them in unsigned.
// 32x16 "high" products map to a single IMAD.HI: (a*b16)>>16 == mulhi(a, b16<<16).
#pragma once
#include <cstdint>

namespace xb {

typedef int32_t i32;
typedef int16_t i16;
typedef int64_t i64;
typedef uint32_t u32;

#define XB_DEV __device__ __forceinline__

XB_DEV i32 wadd(i32 a, i32 b) { return (i32)((u32)a + (u32)b); }
XB_DEV i32 wsub(i32 a, i32 b) { return (i32)((u32)a - (u32)b); }
XB_DEV i32 wneg(i32 a) { return (i32)(0u - (u32)a); }
XB_DEV i32 lsl(i32 a, int s) { return (i32)((u32)a << s); }

// ixheaac_add32_sat / sub32_sat (ops32.h:197,225) == PTX add.sat.s32 / sub.sat.s32
XB_DEV i32 add_sat(i32 a, i32 b) {
  i32 r;
  asm("add.sat.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
XB_DEV i32 sub_sat(i32 a, i32 b) {
  i32 r;
  asm("sub.sat.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// ixheaac_negate32_sat (ops32.h:317)
XB_DEV i32 neg_sat(i32 a) { return sub_sat(0, a); }
// ixheaac_abs32_nrm (ops32.h:283)
XB_DEV i32 abs_nrm(i32 a) { return a ^ (a >> 31); }
// ixheaac_norm32 (ops32.h:236): 31 for 0 and -1
XB_DEV int norm32(i32 a) { return __clz(a ^ (a >> 31)) - 1; }
// ixheaac_pnorm32 (ops32.h:257), argument non-negative
XB_DEV int pnorm32(i32 a) { return __clz(a) - 1; }

// ixheaac_mult32x16in32 (ops40.h:34): b16 is a sign-extended 16-bit value
XB_DEV i32 mul32x16(i32 a, i32 b16) { return __mulhi(a, b16 << 16); }
// multiplier taken from the low / high half of a packed word (aac_imdct.c:80, ops32.h:134)
XB_DEV i32 mul32x16l(i32 a, i32 w) { return __mulhi(a, (i32)((u32)w << 16)); }
XB_DEV i32 mul32x16h(i32 a, i32 w) { return __mulhi(a, (i32)((u32)w & 0xffff0000u)); }
// ixheaac_mult32x16in32_sat (ops32.h:144) / ixheaacd_mult32x16lin32_sat (aac_imdct.c:95): full product, saturated
XB_DEV i32 mul32x16_fullsat(i32 a, i32 b16) {
  i32 lo = a * b16;
  i32 hi = __mulhi(a, b16);
  if (hi != (lo >> 31)) lo = (hi < 0) ? (i32)0x80000000 : 0x7fffffff;
  return lo;
}
// ixheaac_mult32 (ops40.h:78), mult32_shl (:68)
XB_DEV i32 mul32(i32 a, i32 b) { return __mulhi(a, b); }
XB_DEV i32 mul32_shl(i32 a, i32 b) { return lsl(__mulhi(a, b), 1); }

// ixheaac_shl32 (ops32.h:39) / shr32 (:51): count masked to 8 bits
XB_DEV i32 shl32(i32 a, int b) {
  b &= 0xff;
  return b > 31 ? 0 : lsl(a, b);
}
XB_DEV i32 shr32(i32 a, int b) {
  b &= 0xff;
  return a >> (b > 31 ? 31 : b);  // b>=31 -> sign fill, identical to a>>31
}
// ixheaac_shl32_sat (ops32.h:67), 0 <= b <= 31
XB_DEV i32 shl32_sat(i32 a, int b) {
  i32 r = lsl(a, b);
  if ((r >> b) != a) r = (a < 0) ? (i32)0x80000000 : 0x7fffffff;
  return r;
}
// ixheaac_shr32_sat (ops32.h:377): rounding shift
XB_DEV i32 shr32_sat(i32 a, int b) {
  b &= 0xff;
  if (b >= 31) return a >> 31;
  if (b <= 0) return a;
  return add_sat(a, 1 << (b - 1)) >> b;
}
// ixheaac_shl32_dir_sat_limit (ops.h:114)
XB_DEV i32 shl32_dir_sat_limit(i32 a, int b) {
  if (b < 0) {
    b = -b;
    return a >> (b > 31 ? 31 : b);
  }
  return shl32_sat(a, b);
}
// ops16.h
XB_DEV i32 sat16(i32 v) { return max(-32768, min(32767, v)); }
XB_DEV i32 neg16(i32 a) { return a == -32768 ? 32767 : -a; }
XB_DEV i32 round16(i32 a) { return add_sat(a, 0x8000) >> 16; }  // ops16.h:231
XB_DEV i32 sext16(i32 a) { return (i32)(i16)a; }

}  // namespace xb
