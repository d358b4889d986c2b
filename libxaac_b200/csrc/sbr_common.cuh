// sbr_common.cuh — warp-cooperative block-floating-point helpers shared by the SBR stage kernels.
#pragma once
#include "fixmath.cuh"

namespace xb {

// headroom of [slot range] x [band range] (env_calc.c:1159-1207), warp-cooperative
XB_DEV int warp_headroom(const i32 *mat, int b0, int b1, int s0, int s1, int lane) {
  i32 mx = 1;
  const int nb = b1 - b0;
  if (nb > 0) {
    const int total = (s1 - s0) * nb;
    for (int i = lane; i < total; i += 32) {
      const int l = s0 + i / nb, k = b0 + i % nb;
      mx |= abs_nrm(mat[128 * l + k]) | abs_nrm(mat[128 * l + 64 + k]);
    }
  }
  mx = __reduce_or_sync(0xffffffffu, (unsigned)mx);
  return pnorm32(mx);
}

// env_calc.c:1099-1157 (complex), warp-cooperative
XB_DEV void warp_adjust_scale(i32 *mat, int b0, int b1, int s0, int s1, int shift, int lane) {
  if (shift == 0 || b1 <= b0) return;
  shift = max(-31, min(31, shift));
  const int nb = b1 - b0, total = (s1 - s0) * nb;
  for (int i = lane; i < total; i += 32) {
    const int l = s0 + i / nb, k = b0 + i % nb;
    i32 *pr = mat + 128 * l + k, *pi = pr + 64;
    const i32 a = *pr, b = *pi;
    *pr = shift > 0 ? lsl(a, shift) : (a >> -shift);
    *pi = shift > 0 ? lsl(b, shift) : (b >> -shift);
  }
}

}  // namespace xb
