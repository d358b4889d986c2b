/*
 * oracle/src/usac_fd.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain-C restatement of the USAC frequency-domain core transform of libxaac for pure FD streams
 * (previous frame FD, no FAC data, no error concealment, 1024-sample core frames): ixheaacd_fd_frm_dec
 * (decoder/ixheaacd_imdct.c:596) with ixheaacd_fd_imdct_long (:477) / ixheaacd_fd_imdct_short (:336),
 * ixheaacd_acelp_imdct (:186), ixheaacd_fft_based_imdct (:149), the pre / post twiddles (:111, :129), the saturating
 * radix-4 (+ radix-2) FFT ixheaacd_complex_fft_p2_dec (decoder/ixheaacd_fft.c:1412, fft_mode = 1) and the windowing /
 * scaling leaves of decoder/ixheaacd_basic_ops.c.  Pointer walks are restated with explicit indices; every function
 * cites the reference lines it follows.  Pinned against the compiled reference (oracle/_ref: ref_usac_complex_fft,
 * ref_usac_fd_frm_dec) by tests/test_oracle_usac.py.
 */
#include <string.h>
#include "fixmath.h"
#include "xaac_oracle.h"

#define U32T(off) ((const i32 *)(urom + (off)))

static inline i32 mult32_sat(i32 a, i32 b) { return ox_sat64(((i64)a * (i64)b) >> 31); } /* fft.c:48 */
static inline i32 mult32_sh1(i32 a, i32 b) { return (i32)(((i64)a * (i64)b) >> 31); }    /* vec_baisc_ops.h:28 */
static inline i32 ADD(i32 a, i32 b) { return ox_add_sat(a, b); }
static inline i32 SUB(i32 a, i32 b) { return ox_sub_sat(a, b); }
static inline i32 SH1(i32 a) { return ox_shl32_sat(a, 1); }
static inline i32 NEGW(i32 a) { return (i32)(0u - (u32)a); } /* plain unary minus under -fwrapv */

/* rotation forms of the twiddle multiplications (fft.c:2099-2418) */
static inline void rot_a(i32 *xr, i32 *xi, i32 wh, i32 wl) { /* (xr wl + xi wh, -xr wh + xi wl) */
  i32 t = ADD(mult32_sat(*xr, wl), mult32_sat(*xi, wh));
  *xi = ADD(NEGW(mult32_sat(*xr, wh)), mult32_sat(*xi, wl));
  *xr = t;
}
static inline void rot_b(i32 *xr, i32 *xi, i32 wh, i32 wl) { /* (xr wh - xi wl, xr wl + xi wh) */
  i32 t = SUB(mult32_sat(*xr, wh), mult32_sat(*xi, wl));
  *xi = ADD(mult32_sat(*xr, wl), mult32_sat(*xi, wh));
  *xr = t;
}
static inline void rot_c(i32 *xr, i32 *xi, i32 wh, i32 wl) { /* (-(xr wl + xi wh), -xr wh + xi wl) */
  i32 t = NEGW(ADD(mult32_sat(*xr, wl), mult32_sat(*xi, wh)));
  *xi = ADD(NEGW(mult32_sat(*xr, wh)), mult32_sat(*xi, wl));
  *xr = t;
}

/* the radix-4 butterfly shared by every stage (fft.c:1995-2012); alt = 1: the variant of the last twiddle segment
 * (fft.c:2390-2407).  p[0..3] are the four legs (re, im); outputs in the reference's store order. */
static inline void bfly4(i32 *p0, i32 *p1, i32 *p2, i32 *p3, i32 x1r, i32 x1i, i32 x2r, i32 x2i, i32 x3r, i32 x3i, int alt) {
  i32 x0r = p0[0], x0i = p0[1];
  x0r = ADD(x0r, x2r); x0i = ADD(x0i, x2i);
  x2r = SUB(x0r, SH1(x2r)); x2i = SUB(x0i, SH1(x2i));
  x1r = ADD(x1r, x3r);
  if (!alt) { x1i = ADD(x1i, x3i); x3r = SUB(x1r, SH1(x3r)); x3i = SUB(x1i, SH1(x3i)); }
  else { x1i = SUB(x1i, x3i); x3r = SUB(x1r, SH1(x3r)); x3i = ADD(x1i, SH1(x3i)); }
  x0r = ADD(x0r, x1r); x0i = ADD(x0i, x1i);
  x1r = SUB(x0r, SH1(x1r)); x1i = SUB(x0i, SH1(x1i));
  x2r = SUB(x2r, x3i); x2i = ADD(x2i, x3r);
  x3i = ADD(x2r, SH1(x3i)); x3r = SUB(x2i, SH1(x3r));
  p0[0] = x0r; p0[1] = x0i;
  p1[0] = x2r; p1[1] = x2i;
  p2[0] = x1r; p2[1] = x1i;
  p3[0] = x3i; p3[1] = x3r;
}

static inline u32 dig_rev(u32 v, int m) { /* fft.c:39-46 */
  v = ((v & 0x33333333u) << 2) | ((v & ~0x33333333u) >> 2);
  v = ((v & 0x0F0F0F0Fu) << 4) | ((v & ~0x0F0F0F0Fu) >> 4);
  v = ((v & 0x00FF00FFu) << 8) | ((v & ~0x00FF00FFu) >> 8);
  return v >> m;
}

/* decoder/ixheaacd_fft.c:1412-2491 with fft_mode = 1 (npoints = 512 or 64).  Returns the updated preshift. */
int xo_usac_complex_fft(const uint8_t *urom, i32 *xr, i32 *xi, int npoints, int preshift) {
  const i32 *tw = U32T(XO_UROM_FFT_TW);
  i32 px[1024], y[1024];
  const int dig_rev_shift = ox_norm32(npoints) + 1 - 16;
  int n_stages = 30 - ox_norm32(npoints);
  const int not_power_4 = n_stages & 1;
  n_stages >>= 1;
  int n = 0;
  for (int t = npoints; t >> 1; t >>= 1) n++;
  int shift = (n % 2 == 0) ? (n + 4) / 2 : (n + 3) / 2;
  for (int i = 0; i < npoints; i++) { /* :1443-1446, C division: truncation toward zero */
    px[2 * i] = xr[i] / (1 << shift);
    px[2 * i + 1] = xi[i] / (1 << shift);
  }
  for (int i = 0; i < npoints; i += 4) { /* :1969-2020 first radix-4 stage with digit reversal */
    int h2 = (int)dig_rev((u32)i, dig_rev_shift);
    if (not_power_4) { h2 += 1; h2 &= ~1; }
    const i32 *a = px + h2, *b = a + (npoints >> 1), *c = b + (npoints >> 1), *d = c + (npoints >> 1);
    i32 *o = y + 2 * i;
    o[0] = a[0]; o[1] = a[1];
    bfly4(o, o + 2, o + 4, o + 6, b[0], b[1], c[0], c[1], d[0], d[1], 0);
  }
  int del = 4, nodespacing = 64, in_loop_cnt = npoints >> 4;
  for (int st = n_stages - 1; st > 0; st--) { /* :2025-2422 */
    const int sec = 85 * (nodespacing * del) / 256; /* S/4 + S/8 - S/16 + S/32 - S/64 + S/128 - S/256, S = 256 */
    for (int jj = 0; jj < del; jj++) {
      const int j = jj * nodespacing;
      i32 w1h = 0, w1l = 0, w2h = 0, w2l = 0, w3h = 0, w3l = 0;
      int seg = 0;
      if (jj > 0) {
        w1h = tw[2 * j]; w1l = tw[2 * j + 1];
        if (j <= sec) { seg = 1; w2h = tw[4 * j]; w2l = tw[4 * j + 1]; w3h = tw[6 * j]; w3l = tw[6 * j + 1]; }
        else if (j <= (nodespacing * del) >> 1) { seg = 2; w2h = tw[4 * j]; w2l = tw[4 * j + 1]; w3h = tw[6 * j - 512]; w3l = tw[6 * j - 511]; }
        else if (j <= sec * 2) { seg = 3; w2h = tw[4 * j - 512]; w2l = tw[4 * j - 511]; w3h = tw[6 * j - 512]; w3l = tw[6 * j - 511]; }
        else { seg = 4; w2h = tw[4 * j - 512]; w2l = tw[4 * j - 511]; w3h = tw[6 * j - 1024]; w3l = tw[6 * j - 1023]; }
      }
      for (int g = 0; g < in_loop_cnt; g++) {
        i32 *p0 = y + 2 * jj + 8 * del * g, *p1 = p0 + 2 * del, *p2 = p1 + 2 * del, *p3 = p2 + 2 * del;
        i32 x1r = p1[0], x1i = p1[1], x2r = p2[0], x2i = p2[1], x3r = p3[0], x3i = p3[1];
        if (seg) {
          rot_a(&x1r, &x1i, w1h, w1l);
          if (seg <= 2) rot_a(&x2r, &x2i, w2h, w2l); else rot_b(&x2r, &x2i, w2h, w2l);
          if (seg == 1) rot_a(&x3r, &x3i, w3h, w3l);
          else if (seg <= 3) rot_b(&x3r, &x3i, w3h, w3l);
          else rot_c(&x3r, &x3i, w3h, w3l);
        }
        bfly4(p0, p1, p2, p3, x1r, x1i, x2r, x2i, x3r, x3i, seg == 4);
      }
    }
    nodespacing >>= 2;
    del <<= 2;
    in_loop_cnt >>= 2;
  }
  if (not_power_4) { /* :2423-2481 final radix-2 stage */
    nodespacing <<= 1;
    shift += 1;
    for (int half = 0; half < 2; half++)
      for (int t = 0; t < del / 2; t++) {
        const i32 wh = tw[2 * nodespacing * t], wl = tw[2 * nodespacing * t + 1];
        i32 *p0 = y + 2 * (half * (del / 2) + t), *p1 = p0 + 2 * del;
        i32 x0r = p0[0], x0i = p0[1], x1r = p1[0], x1i = p1[1];
        if (!half) rot_a(&x1r, &x1i, wh, wl); else rot_b(&x1r, &x1i, wh, wl);
        p1[0] = ox_sub(x0r / 2, x1r / 2);
        p1[1] = ox_sub(x0i / 2, x1i / 2);
        p0[0] = ox_add(x0r / 2, x1r / 2);
        p0[1] = ox_add(x0i / 2, x1i / 2);
      }
  }
  for (int i = 0; i < npoints; i++) { xr[i] = y[2 * i]; xi[i] = y[2 * i + 1]; }
  return shift - preshift;
}

/* decoder/ixheaacd_imdct.c:186-209 + :149-184, :111-147 for npoints = 2 * N, N = 1024 or 128 (power of two only).
 * x [N] in place; returns the updated qshift. */
static int acelp_imdct(const uint8_t *urom, i32 *x, int N, int qshift) {
  int preshift = 0;
  for (int k = N; ((k & 1) == 0) && k != 1; k >>= 1) preshift++;
  const int nl = N >> 1;
  const i32 *cs = U32T(nl == 512 ? XO_UROM_COS512 : XO_UROM_COS64), *sn = U32T(nl == 512 ? XO_UROM_SIN512 : XO_UROM_SIN64);
  i32 r[512], im[512];
  for (int i = 0; i < nl; i++) { /* :111-127 */
    const i32 a = x[2 * i], b = x[2 * nl - 1 - 2 * i];
    r[i] = ox_sub(ox_mul32(ox_neg_sat(a), cs[i]), ox_mul32(b, sn[i]));
    im[i] = ox_sub(ox_mul32(b, cs[i]), ox_mul32(a, sn[i]));
  }
  preshift = xo_usac_complex_fft(urom, r, im, nl, preshift);
  for (int i = 0; i < nl; i++) { /* :129-147 */
    x[2 * i] = NEGW(ox_sub(ox_mul32(r[i], cs[i]), ox_mul32(im[i], sn[i])));
    x[2 * nl - 1 - 2 * i] = NEGW(ox_add(ox_mul32(im[i], cs[i]), ox_mul32(r[i], sn[i])));
  }
  preshift += 2;
  return (int8_t)(qshift - preshift);
}

static int calc_max_spectralline(const i32 *p, int n) { /* imdct.c:81-91 */
  i32 m = 0;
  for (int k = 0; k < n; k++) {
    const i32 a = p[k] == OX_MIN32 ? OX_MAX32 : (p[k] < 0 ? -p[k] : p[k]);
    if (a > m) m = a;
  }
  return ox_norm32(m);
}
/* imdct.c:93-99.  The reference calls this with max_shift - 1, i.e. -1 when the transform output has no headroom left
 * (only pathological spectra: full-scale alternating, DC).  `x << -1` is undefined in ISO C; the reference build this
 * oracle is pinned to (gcc -O3, x86-64) vectorises the loop with PSLLD, which yields 0 for every count > 31. */
static void normalize(i32 *b, int shift, int n) {
  for (int i = 0; i < n; i++) b[i] = shift < 0 ? 0 : ox_lsl(b[i], shift);
}
static const i32 *window(const uint8_t *urom, int len, int sel) { /* ixheaacd_calc_window, Windowing.c:29-111 */
  if (len == 1024) return U32T(sel ? XO_UROM_KBD1024 : XO_UROM_SINE1024);
  return U32T(sel ? XO_UROM_KBD128 : XO_UROM_SINE128);
}
static void scale_down(i32 *d, const i32 *s, int len, int s1, int s2) { /* basic_ops.c:623-639 */
  for (int i = 0; i < len; i++) d[i] = s1 > s2 ? (s[i] >> (s1 - s2)) : ox_shl32_sat(s[i], s2 - s1);
}

/* basic_ops.c:77-123 */
static int windowing_long1(const i32 *src1, const i32 *src2, const i32 *win, i32 *dest, int vlen, int s1, int s2) {
  for (int i = 0; i < vlen / 2; i++) {
    const i32 wf = win[i], wr = win[vlen - 1 - i], a = src1[i], ov = src2[i], ovr = src2[vlen - 1 - i];
    if (s1 > s2) {
      dest[i] = ADD(mult32_sh1(a, wf) >> (s1 - s2), mult32_sh1(ov, wr));
      dest[vlen - 1 - i] = ADD(mult32_sh1(ox_neg_sat(a), wr) >> (s1 - s2), mult32_sh1(ovr, wf));
    } else {
      dest[i] = ADD(mult32_sh1(a, wf), mult32_sh1(ov, wr) >> (s2 - s1));
      dest[vlen - 1 - i] = ADD(mult32_sh1(ox_neg_sat(a), wr), mult32_sh1(ovr, wf) >> (s2 - s1));
    }
  }
  return s1 > s2 ? s2 : s1;
}

/* basic_ops.c:298-372 (no FAC): n_flat = 448, n_trans = 128, n_long = 1024; src1 = spectrum + 512 */
static int windowing_long3(const i32 *src1, const i32 *wsh, const i32 *ov, i32 *dest, int shiftp, int so) {
  const int nf = 448, nt = 128, nl = 1024;
  for (int i = 0; i < nl; i++) {
    if (shiftp > so) {
      if (i < nf) dest[i] = ov[i];
      else if (i < nl / 2) dest[i] = ADD(mult32_sh1(src1[i], wsh[i - nf]) >> (shiftp - so), mult32_sh1(ov[i], wsh[nt - 1 - (i - nf)]));
      else if (i < nf + nt) dest[i] = ADD(mult32_sh1(ox_neg_sat(src1[nl - i - 1]), wsh[i - nf]) >> (shiftp - so), mult32_sh1(ov[i], wsh[nt - 1 - (i - nf)]));
      else dest[i] = ox_neg_sat(src1[nl - i - 1]) >> (shiftp - so);
    } else {
      if (i < nf) dest[i] = ov[i] >> (so - shiftp);
      else if (i < nl / 2) dest[i] = ADD(mult32_sh1(src1[i], wsh[i - nf]), mult32_sh1(ov[i], wsh[nt - 1 - (i - nf)]) >> (so - shiftp));
      else if (i < nf + nt) dest[i] = ADD(mult32_sh1(ox_neg_sat(src1[nl - i - 1]), wsh[i - nf]), mult32_sh1(ov[i], wsh[nt - 1 - (i - nf)]) >> (so - shiftp));
      else dest[i] = ox_neg_sat(src1[nl - i - 1]);
    }
  }
  return shiftp > so ? so : shiftp;
}

/* decoder/ixheaacd_imdct.c:477-594 (td_frame_prev = 0, fac_apply = 0) */
static int fd_imdct_long(const uint8_t *urom, i32 *in, i32 *ov, int win_seq, int shape_prev, i32 *out) {
  const int so = 14;
  int max_shift = calc_max_spectralline(in, 1024);
  normalize(in, max_shift, 1024);
  int shiftp = (int8_t)(max_shift + 6);
  shiftp = acelp_imdct(urom, in, 1024, shiftp);
  max_shift = calc_max_spectralline(in, 1024);
  normalize(in, max_shift - 1, 1024);
  shiftp = (int8_t)(shiftp + max_shift - 1);
  if (shiftp - so > 31) shiftp = 31 + so;
  int output_q = 0;
  switch (win_seq) {
    case 0: case 1: /* ONLY_LONG, LONG_START */
      output_q = windowing_long1(in + 512, ov, window(urom, 1024, shape_prev), out, 1024, shiftp, so);
      break;
    case 3: case 4: /* LONG_STOP, STOP_START */
      output_q = windowing_long3(in + 512, window(urom, 128, shape_prev), ov, out, shiftp, so);
      break;
    default: break;
  }
  for (int i = 0; i < 512; i++) { /* :553-568: a right shift on both sides of the comparison */
    const i32 v = ox_neg_sat(in[i]) >> (shiftp > so ? shiftp - so : so - shiftp);
    ov[512 + i] = v;
    ov[511 - i] = v;
  }
  for (int i = 0; i < 1024; i++) /* ixheaacd_scale_down_adj(.., output_q, 15), basic_ops.c:641-657, ADJ_SCALE = 11 */
    out[i] = ADD(output_q > 15 ? (out[i] >> (output_q - 15)) : ox_shl32_sat(out[i], 15 - output_q), 11);
  return 0;
}

/* basic_ops.c:429-478: src1 = spectrum + 64, fp = overlap work buffer + 448 */
static void windowing_short2(const i32 *src1, const i32 *win, i32 *fp, int shiftp, int so) {
  const int ns = 128, nf = 448;
  for (int i = 0; i < ns / 2; i++) {
    const i32 wf = win[i], wr = win[ns - 1 - i];
    if (so > shiftp) {
      fp[i] = ADD(mult32_sh1(src1[i], wf), mult32_sh1(fp[i], wr) >> (so - shiftp));
      fp[ns - i - 1] = ADD(mult32_sh1(ox_neg_sat(src1[i]), wr), mult32_sh1(fp[ns - i - 1], wf) >> (so - shiftp));
    } else {
      fp[i] = ADD(mult32_sh1(src1[i], wf) >> (shiftp - so), mult32_sh1(fp[i], wr));
      fp[ns - i - 1] = ADD(mult32_sh1(ox_neg_sat(src1[i]), wr) >> (shiftp - so), mult32_sh1(fp[ns - i - 1], wf));
    }
  }
  for (int i = ns; i < nf + ns; i++) fp[i] = 0;
}
/* basic_ops.c:480-521: win = short window table; win_rev = win + 127 walks down, win_fwd = win walks up */
static int windowing_short3(const i32 *src1, const i32 *win, i32 *fp, int shiftp, int so) {
  const int ns = 128;
  for (int i = 0; i < ns / 2; i++) {
    const i32 wr = win[ns - 1 - i], wf = win[i], a = ox_neg_sat(src1[ns / 2 - i - 1]);
    if (so > shiftp) {
      fp[i] = ADD(mult32_sh1(a, wr), fp[i] >> (so - shiftp));
      fp[ns - i - 1] = ADD(mult32_sh1(a, wf), fp[ns - i - 1] >> (so - shiftp));
    } else {
      fp[i] = ADD(mult32_sh1(a, wr) >> (shiftp - so), fp[i]);
      fp[ns - i - 1] = ADD(mult32_sh1(a, wf) >> (shiftp - so), fp[ns - i - 1]);
    }
  }
  return so > shiftp ? shiftp : so;
}
/* basic_ops.c:523-621: win_fwd = win[i], win_rev = win[127-i]; win_fwd1 = win + 127 walks down, win_rev1 = win walks up */
static int windowing_short4(const i32 *src1, const i32 *win, i32 *fp, int flag, int shiftp, int so, int oq) {
  const int ns = 128;
  const int big = so > oq;
  const int sh = big ? shiftp - oq : shiftp - so;
  for (int i = 0; i < ns / 2; i++) {
    const i32 wf = win[i], wr = win[ns - 1 - i], a = src1[ns / 2 + i];
    if (big) {
      fp[i] = ADD(mult32_sh1(a, wf) >> sh, fp[i]);
      fp[ns - i - 1] = ADD(mult32_sh1(ox_neg_sat(a), wr) >> sh, fp[ns - i - 1]);
    } else {
      fp[i] = ADD(mult32_sh1(a, wf) >> sh, fp[i] >> (oq - so));
      fp[ns - i - 1] = ADD(mult32_sh1(ox_neg_sat(a), wr) >> sh, fp[ns - i - 1]);
    }
  }
  for (int i = ns / 2; i < ns; i++) {
    const int t = i - ns / 2; /* win_fwd1 = win[127 - t], win_rev1 = win[t] */
    const i32 a = ox_neg_sat(src1[ns - i - 1]);
    i32 *pa = &fp[i + ns / 2], *pb = &fp[3 * ns - ns / 2 - i - 1];
    const i32 va = flag ? mult32_sh1(a, win[ns - 1 - t]) : a, vb = flag ? mult32_sh1(a, win[t]) : a;
    if (big) {
      *pa = ADD(va >> sh, *pa >> (so - oq));
      *pb = ADD(vb >> sh, *pb >> (so - oq));
    } else {
      *pa = ADD(va >> sh, *pa);
      *pb = ADD(vb >> sh, *pb);
    }
  }
  return big ? oq : so;
}

/* decoder/ixheaacd_imdct.c:336-475 (td_frame_prev = 0, fac_apply = 0) */
static int fd_imdct_short(const uint8_t *urom, i32 *in, i32 *ov, int shape, int shape_prev, i32 *out) {
  const int so = 14, ns = 128, nf = 448;
  i32 buf[2048];
  memset(buf, 0, sizeof(buf));
  int max_shift = calc_max_spectralline(in, 1024);
  normalize(in, max_shift, 1024);
  const int input_q = (int8_t)(max_shift + 6);
  int shiftp = input_q;
  memcpy(buf, ov, 1024 * sizeof(i32));
  i32 *fp = buf + nf;
  for (int k = 0; k < 8; k++) shiftp = acelp_imdct(urom, in + k * ns, ns, input_q);
  max_shift = calc_max_spectralline(in, 1024);
  normalize(in, max_shift - 1, 1024);
  shiftp = (int8_t)(shiftp + max_shift - 1);
  if (shiftp - so > 31) shiftp = 31 + so;
  const i32 *wsh = window(urom, 128, shape), *wprev = window(urom, 128, shape_prev);
  windowing_short2(in + ns / 2, wprev, fp, shiftp, so);
  int oq = windowing_short3(in, wsh, fp + ns, shiftp, so);
  const i32 *p = in + ns;
  fp += ns;
  for (int k = 1; k < 7; k++, p += ns, fp += ns) oq = windowing_short4(p, wsh, fp, 1, shiftp, so, oq);
  oq = windowing_short4(p, wsh, fp, 0, shiftp, so, oq);
  memset(buf + 2048 - nf, 0, nf * sizeof(i32));
  scale_down(buf, buf, nf, so, oq);
  scale_down(ov, buf + 1024, 1024, oq, so);
  scale_down(out, buf, 1024, oq, 15);
  return 0;
}

/* decoder/ixheaacd_imdct.c:596-654 for one channel of a pure FD stream: coef [1024] (destroyed), ov [1024] in/out,
 * out [1024].  win_seq: 0 ONLY_LONG, 1 LONG_START, 2 EIGHT_SHORT, 3 LONG_STOP, 4 STOP_START. */
int xo_usac_fd_frm_dec(const uint8_t *urom, i32 *coef, i32 *ov, int win_seq, int win_shape, int win_shape_prev, i32 *out) {
  if (win_seq != 2) return fd_imdct_long(urom, coef, ov, win_seq, win_shape_prev, out);
  return fd_imdct_short(urom, coef, ov, win_shape, win_shape_prev, out);
}
void xo_usac_fd_frm_dec_batch(const uint8_t *urom, i32 *coef, i32 *ov, const i32 *win_seq, const i32 *win_shape,
                              const i32 *win_shape_prev, i32 *out, i32 *err, int n) {
  for (int u = 0; u < n; u++)
    err[u] = xo_usac_fd_frm_dec(urom, coef + (size_t)u * 1024, ov + (size_t)u * 1024, win_seq[u], win_shape[u],
                                win_shape_prev[u], out + (size_t)u * 1024);
}
