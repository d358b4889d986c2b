// aac_spectral_kernel.cu — the AAC pre-IMDCT spectral stage for sm_100a (B200): what ixheaacd_channel_pair_process
// (decoder/ixheaacd_channel.c:602-718) does to the dequantised, scale-factor-applied spectrum of one element for AAC-LC:
//   ixheaacd_ms_stereo_process          decoder/ixheaacd_stereo.c:54-116      (add / sub, saturating, per used sfb)
//   ixheaacd_intensity_stereo_process   decoder/ixheaacd_stereo.c:129-236     (right = scaled left, per intensity sfb)
//   ixheaacd_aac_tns_process            decoder/ixheaacd_pns_js_thumb.c:248-514 with the selector leaves
//     ixheaacd_tns_decode_coef (:202), ixheaacd_tns_parcor_lpc_convert_dec (decoder/ixheaacd_aac_tns.c:147),
//     ixheaacd_calc_max_spectral_line_dec (:422), ixheaacd_tns_ar_filter_dec (:371)
//   ixheaacd_map_ms_mask_pns (channel.c:703-726), ixheaacd_pns_process / ixheaacd_gen_rand_vec (pns_js_thumb.c:74-200) with
//     ixheaacd_sqrt / ixheaacd_one_by_sqrt_calc (decoder/ixheaacd_basic_funcs.c:155-196), ixheaac_div32_pos_normb
// Three kernels: the stereo tools with a warp per element (rows of an sfb are coalesced 128-byte requests); perceptual noise
// substitution with a THREAD per element (one linear-congruential generator runs through all noise bands of both channels, and
// on into the next frame); TNS with a thread per channel — the all-pole filter is a recursion over up to 1024 spectral lines
// whose saturating accumulation fixes the order of every add.  For the serial parts the parallelism is across the batch.
// Record layout: XAAC_SPS_* of include/xaac_b200.h (the reference's own structs, byte for byte, where they are plain data).
#include <cstdint>
#include <cuda_runtime.h>
#include "fixmath.cuh"
#include "kernels.h"

namespace xb {
namespace {

struct ChanView {
  const unsigned char *b;
  XB_DEV int word(int i) const { return reinterpret_cast<const int *>(b)[i]; }
  XB_DEV int window_sequence() const { return word(0); }
  XB_DEV int max_sfb() const { return word(1); }
  XB_DEV int num_window_groups() const { return word(2); }
  XB_DEV int pns_active() const { return word(3); }
  XB_DEV int tns_max_bands() const { return word(4); }
  XB_DEV int group_len(int g) const { return (signed char)b[kSpsChGroupLen + g]; }
  XB_DEV int code_book(int i) const { return (signed char)b[kSpsChCodeBook + i]; }
  XB_DEV int scale_factor(int i) const { return reinterpret_cast<const int16_t *>(b + kSpsChScaleFactor)[i]; }
  XB_DEV const unsigned char *tns() const { return b + kSpsChTns; }
  XB_DEV int sfb_index(int i) const { return reinterpret_cast<const int16_t *>(b + kSpsChSfbIndex)[i]; }
  XB_DEV int pns_used(int i) const { return b[kSpsChPnsUsed + i]; }
};

// Everything ixheaacd_aac_tns_process decides from the side information alone about one filter (no spectral data involved).
struct TnsPlan {
  int apply;      // 0: the reference skips it
  int start, size, position, order, order_r, dir, res;
  int first, last;  // spectral lines the AR filter touches, relative to the window's first line (reference quirk: max(order_r, size))
};
XB_DEV TnsPlan tns_plan(const ChanView &c, int win, int filt) {
  TnsPlan t;
  t.apply = 0;
  const unsigned char *f = c.tns() + 12 + (win * 3 + filt) * 38;
  t.order = (signed char)f[6];
  t.dir = (signed char)f[4];
  t.res = (signed char)f[5];
  t.start = t.size = t.position = t.order_r = t.first = t.last = 0;
  if (t.order <= 0) return t;
  const int start_band = reinterpret_cast<const int16_t *>(f)[0], stop_band = reinterpret_cast<const int16_t *>(f)[1];
  const int lim = min(c.tns_max_bands(), c.max_sfb());
  const int sb = min(start_band, lim), eb = min(stop_band, lim);
  if (sb < 0 || eb < 0 || sb > 51 || eb > 51) { t.apply = -1; return t; }
  const int start = c.sfb_index(sb), stop = c.sfb_index(eb);
  t.start = start;
  t.size = stop - start;
  if (t.size <= 0) return t;
  const int base = win << 7;
  if (t.dir == -1) {
    t.position = stop - 1;
    if (base + t.position < t.order) return t;
  } else {
    t.position = start;
    if (base + t.position + t.order > 1024) return t;
  }
  // ixheaacd_tns_ar_filter_dec pads the order to a multiple of four (aac_tns.c:378-388) and always runs `order` steps first
  int orr = t.order;
  if (orr & 3) orr = ((orr & ~3) + 4 < 32) ? (orr & ~3) + 4 : 31;
  t.order_r = orr;
  const int steps = max(orr, t.size);
  if (t.dir == -1) { t.last = t.position; t.first = t.position - (steps - 1); }
  else { t.first = t.position; t.last = t.position + steps - 1; }
  t.apply = 1;
  return t;
}

}  // namespace

// ---- kernel 1: validation, M/S, intensity stereo; a warp per element ----
__global__ void __launch_bounds__(256) aac_stereo_tools_kernel(const AacSpectralArgs p) {
  const int lane = threadIdx.x & 31;
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < p.n_units; u += warps_total) {
    const unsigned char *rec = p.side + u * kSpsBytes;
    const int *hdr = reinterpret_cast<const int *>(rec);
    const int num_ch = hdr[kSpsNumCh], common_window = hdr[kSpsCommonWindow];
    ChanView ch[2] = {{rec + kSpsCh}, {rec + kSpsCh + kSpsChBytes}};
    int bad = (num_ch < 1 || num_ch > 2) ? 1 : 0;
    for (int c = 0; c < 2 && !bad; c++) {
      if (c >= num_ch) break;
      const int ws = ch[c].window_sequence(), ms = ch[c].max_sfb(), ng = ch[c].num_window_groups();
      if (ws < 0 || ws > 3 || ms < 0 || ms > (ws == 2 ? 15 : 51) || ng < 1 || ng > 8 || (ch[c].pns_active() && !p.pns_seed)) bad = 1;
      int tot = 0;
      for (int g = 0; g < ng && !bad; g++) {
        const int gl = ch[c].group_len(g);
        if (gl < 1) bad = 1;
        tot += gl;
      }
      if (!bad && tot != (ws == 2 ? 8 : 1)) bad = 1;
      if (!bad && reinterpret_cast<const int *>(ch[c].tns())[0]) {  // tns_data_present: every filter must stay inside the window set
        const int nwin = ws == 2 ? 8 : 1;
        for (int e = lane; e < nwin * 3; e += 32) {
          const int win = e / 3, filt = e - win * 3;
          const int nf = (signed char)ch[c].tns()[4 + win];
          if (nf < 0 || nf > 3) bad = 1;
          else if (filt < nf) {
            const TnsPlan t = tns_plan(ch[c], win, filt);
            if (t.apply < 0 || t.order > 31) bad = 1;
            if (t.apply > 0 && ((win << 7) + t.first < 0 || (win << 7) + t.last > 1023)) bad = 1;
          }
        }
        bad = __any_sync(0xffffffffu, bad) ? 1 : 0;
      }
    }
    if (p.err && lane == 0) p.err[u] = bad ? -2 : 0;
    if (bad || num_ch < 2) continue;
    int32_t *l_spec = p.spec + u * 2048, *r_spec = l_spec + 1024;
    const unsigned char *ms_used = rec + kSpsMsUsed;
    // ---- ixheaacd_ms_stereo_process: groups / lengths / max_sfb of LEFT, band widths of RIGHT's window sequence
    // ixheaacd_map_ms_mask_pns first takes the bands that are noise in BOTH channels out of the mask (channel.c:703-726)
    const bool map_pns = common_window && (ch[0].pns_active() || ch[1].pns_active());
    if (common_window) {
      const int max_sfb = ch[0].max_sfb();
      int w = 0;
      for (int g = 0; g < ch[0].num_window_groups(); g++)
        for (int gl = 0; gl < ch[0].group_len(g); gl++, w++) {
          const int base = w << 7;
          for (int sfb = 0; sfb < max_sfb; sfb++) {
            if (!ms_used[g * 64 + sfb]) continue;
            if (map_pns && ch[0].pns_used((g << 4) + sfb) && ch[1].pns_used((g << 4) + sfb)) continue;
            const int k0 = ch[1].sfb_index(sfb), k1 = ch[1].sfb_index(sfb + 1);
            for (int k = k0 + lane; k < k1 && base + k < 1024; k += 32) {
              const i32 a = l_spec[base + k], b = r_spec[base + k];
              l_spec[base + k] = add_sat(a, b);
              r_spec[base + k] = sub_sat(a, b);
            }
          }
        }
      __syncwarp();
    }
    // ---- ixheaacd_intensity_stereo_process (AAC-LC: code book >= INTENSITY_HCB2), everything from RIGHT
    {
      const int max_sfb = ch[1].max_sfb();
      int w = 0;
      for (int g = 0; g < ch[1].num_window_groups(); g++)
        for (int gl = 0; gl < ch[1].group_len(g); gl++, w++) {
          const int base = w << 7;
          for (int sfb = 0; sfb < max_sfb; sfb++) {
            const int cb = ch[1].code_book(16 * g + sfb);
            if (cb < 14) continue;
            const int sfb_factor = ch[1].scale_factor(16 * g + sfb);
            int scf_exp = sfb_factor >> 2;
            i32 scale = __ldg(p.rom + kBromScaleTable + (sfb_factor & 3));
            int msu = ms_used[g * 64 + sfb] ? 1 : 0;
            if (msu && map_pns && g < ch[0].num_window_groups() && sfb < ch[0].max_sfb() && ch[0].pns_used((g << 4) + sfb) &&
                ch[1].pns_used((g << 4) + sfb))
              msu = 0;
            if (!(msu ^ (cb & 1))) scale = wneg(scale);
            scf_exp = -(scf_exp + 2);
            const int k0 = ch[1].sfb_index(sfb), k1 = ch[1].sfb_index(sfb + 1);
            for (int k = k0 + lane; k < k1 && base + k < 1024; k += 32) {
              i32 t = l_spec[base + k];
              int sh = norm32(t);
              t = shl32(t, sh);
              t = (i32)(((long long)t * (long long)scale) >> 16);
              sh += scf_exp;
              if (sh < 0) t = shl32_sat(t, min(31, -sh));
              else t = shr32(t, min(31, sh));
              r_spec[base + k] = t;
            }
          }
        }
    }
    __syncwarp();
  }
}

// ---- kernel 2: perceptual noise substitution, a thread per element ----
namespace {
XB_DEV i32 mult32_shl_sat(i32 a, i32 b) {
  if (a == (i32)0x80000000 && b == (i32)0x80000000) return 0x7fffffff;
  return lsl(__mulhi(a, b), 1);
}
XB_DEV i32 mult32x16_shl(i32 a, i32 b16) { return lsl(mul32x16(a, b16), 1); }
XB_DEV i32 shl32_dir_sat_limit_(i32 a, int b) { return b < 0 ? shr32(a, min(-b, 31)) : shl32_sat(a, b); }
XB_DEV i32 shr32_dir_sat_limit_(i32 a, int b) { return b < 0 ? shl32_sat(a, -b) : shr32(a, min(b, 31)); }
// decoder/ixheaacd_basic_funcs.c:155-181
XB_DEV i32 one_by_sqrt_calc(i32 op) {
  i32 a = add_sat((i32)0x900ebee0, mult32x16_shl(op, 0x39d9));
  // ixheaac_mult32x16h_in32_shl_sat(op, a): its saturation test compares the 32-bit b with (WORD16)0x8000
  i32 iy = add_sat(0x573b645a, (op == (i32)0x80000000 && a == -32768) ? 0x7fffffff : mult32x16_shl(op, a >> 16));
  iy = shl32_dir_sat_limit_(iy, 1);
#pragma unroll
  for (int r = 0; r < 3; r++) {
    a = mult32_shl_sat(op, iy);
    a = sub_sat(0x40000000, shl32_dir_sat_limit_(mult32_shl_sat(a, iy), 1));
    iy = add_sat(iy, mult32_shl_sat(a, iy));
  }
  return iy;
}
XB_DEV i32 fix_sqrt(i32 op) {  // ixheaacd_sqrt, :183-196
  if (op == 0) return 0;
  int shift = (int16_t)(norm32(op) & ~1);
  op = shl32_dir_sat_limit_(op, shift);
  shift = shr32_dir_sat_limit_(shift, 1);
  op = mult32_shl_sat(one_by_sqrt_calc(op), op);
  return shr32_dir_sat_limit_(op, sat16(shift - 1));
}
XB_DEV i32 div32_pos_normb(i32 a, i32 b) {  // common/ixheaac_basic_ops.h:74-99
  if (a == b) return 0x7fffffff;
  u32 nr = (u32)a, dr = (u32)b, q = 0;
  for (int i = 0; i < 32; i++) {
    q <<= 1;
    if (nr >= dr) { nr -= dr; q += 1; }
    nr <<= 1;
  }
  return (i32)q;
}
// ixheaacd_gen_rand_vec (pns_js_thumb.c:74-112): `count` = sfb_width + 1 lines
XB_DEV void gen_rand_vec(i32 scale, int shift, i32 *spec, int count, i32 &seed) {
  i32 nrg = 0;
  for (int i = 0; i < count; i++) {
    seed = (i32)(1664525u * (u32)seed + 1013904223u);
    const i32 v = seed >> 3;
    spec[i] = v;
    nrg = add_sat(nrg, mult32_shl_sat(v, v));
  }
  int nrg_scale = norm32(nrg);
  if (nrg_scale > 0) {
    nrg_scale &= ~1;
    nrg = shl32_sat(nrg, nrg_scale);
    shift = shift - (nrg_scale >> 1);
  }
  nrg = fix_sqrt(nrg);
  scale = div32_pos_normb(scale, nrg);
  if (shift < -31) shift = -31;
  for (int i = 0; i < count; i++) spec[i] = shr32_dir_sat_limit_(mult32_shl_sat(spec[i], scale), shift);
}
}  // namespace

__global__ void __launch_bounds__(128) aac_pns_kernel(const AacSpectralArgs p) {
  const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= p.n_units || !p.pns_seed || p.err[u] != 0) return;
  const unsigned char *rec = p.side + u * kSpsBytes;
  const int num_ch = reinterpret_cast<const int *>(rec)[kSpsNumCh], common_window = reinterpret_cast<const int *>(rec)[kSpsCommonWindow];
  const ChanView ch[2] = {{rec + kSpsCh}, {rec + kSpsCh + kSpsChBytes}};
  const bool any = ch[0].pns_active() || (num_ch > 1 && ch[1].pns_active());
  if (!any) return;
  // correlation flags: what the parser left plus every band of LEFT's mask (ixheaacd_map_ms_mask_pns / ixheaacd_set_corr_info)
  unsigned corr[4];
  for (int i = 0; i < 4; i++) corr[i] = reinterpret_cast<const unsigned *>(rec + kSpsCorrelated)[i];
  if (num_ch > 1 && common_window)
    for (int g = 0; g < ch[0].num_window_groups(); g++)
      for (int sfb = 0; sfb < ch[0].max_sfb(); sfb++)
        if (rec[kSpsMsUsed + g * 64 + sfb]) {
          const int band = (g << 4) + sfb;
          corr[band >> 5] |= 1u << (band & 31);
        }
  i32 seed = p.pns_seed[u];
  i32 rv[128];  // random_vector: written by the first channel, consumed (and advanced) by the second
  for (int i = 0; i < 128; i++) rv[i] = 0;
  for (int c = 0; c < num_ch; c++) {
    if (!ch[c].pns_active()) continue;
    i32 *spec = p.spec + u * 2048 + c * 1024;
    int w = 0;
    for (int g = 0; g < ch[c].num_window_groups(); g++)
      for (int gl = 0; gl < ch[c].group_len(g); gl++, w++) {
        const int base = w << 7;
        for (int sfb = 0; sfb < ch[c].max_sfb(); sfb++) {
          const int band = (g << 4) + sfb;
          if (!ch[c].pns_used(band)) continue;
          const int sf = ch[c].scale_factor(band);
          const i32 scale_mant = __ldg(p.rom + kBromScaleMant + (sf & 3));
          const int scale_exp = (31 - (sf >> 2)) - 4;
          const int k0 = ch[c].sfb_index(sfb), cnt = ch[c].sfb_index(sfb + 1) - k0;
          if (cnt <= 0 || base + k0 + cnt > 1024) continue;
          i32 *ps = spec + base + k0;
          if ((corr[band >> 5] >> (band & 31)) & 1) {
            if (c == 0) {
              rv[band] = seed;
              gen_rand_vec(scale_mant, scale_exp, ps, cnt, seed);
            } else {
              gen_rand_vec(scale_mant, scale_exp, ps, cnt, rv[band]);
            }
          } else {
            gen_rand_vec(scale_mant, scale_exp, ps, cnt, seed);
          }
        }
      }
  }
  p.pns_seed[u] = seed;
}

// ---- kernel 3: TNS, a thread per channel ----
namespace {
XB_DEV i32 mult16x16_shl_sat(i32 a, i32 b) {
  const i32 pr = a * b;
  return pr != (i32)0x40000000 ? lsl(pr, 1) : 0x7fffffff;
}
XB_DEV i32 abs32_sat(i32 a) { return a == (i32)0x80000000 ? 0x7fffffff : (a < 0 ? -a : a); }
}  // namespace

__global__ void __launch_bounds__(128) aac_tns_kernel(const AacSpectralArgs p) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * p.n_units) return;
  const long long u = t >> 1;
  const int c = (int)(t & 1);
  if (p.err && p.err[u] != 0) return;
  const unsigned char *rec = p.side + u * kSpsBytes;
  if (c >= reinterpret_cast<const int *>(rec)[kSpsNumCh]) return;
  const ChanView ch{rec + kSpsCh + c * kSpsChBytes};
  const unsigned char *tns = ch.tns();
  if (!reinterpret_cast<const int *>(tns)[0]) return;
  i32 *spec = p.spec + u * 2048 + c * 1024;
  const int nwin = ch.window_sequence() == 2 ? 8 : 1;
  for (int win = 0; win < nwin; win++) {
    const int nf = (signed char)tns[4 + win];
    for (int filt = 0; filt < nf; filt++) {
      const TnsPlan tp = tns_plan(ch, win, filt);
      if (tp.order <= 0 || tp.size <= 0) continue;
      const unsigned char *f = tns + 12 + (win * 3 + filt) * 38;
      // ixheaacd_tns_decode_coef
      int16_t parcor[32], lpc[33];
      {
        const int off = tp.res ? 8 : 4;
        const int tb = tp.res ? kBromTnsCoeff4 : kBromTnsCoeff3;
        for (int o = 0; o < tp.order; o++) {
          const int idx = (signed char)f[7 + o] + off;
          parcor[o] = reinterpret_cast<const int16_t *>(p.rom)[tb + (idx & (tp.res ? 15 : 7))];
        }
      }
      // ixheaacd_tns_parcor_lpc_convert_dec (runs before the position checks, but has no side effect besides lpc / scale)
      int scale_lpc = 0;
      {
        const int order = tp.order;
        int status = 1;
        while (status) {
          status = 0;
          int16_t b1[32], b2[32];
          for (int i = 0; i < 32; i++) { b1[i] = 0; b2[i] = 0; }
          i32 accu1 = 0x7fffffff >> scale_lpc;
          for (int i = 0; i <= order; i++) {
            const i32 accu = accu1;
            for (int j = 0; j < order; j++) {
              b2[j] = (int16_t)round16(accu1);
              accu1 = add_sat(accu1, mult16x16_shl_sat(parcor[j], b1[j]));
              if (abs32_sat(accu1) == 0x7fffffff) status = 1;
            }
            for (int j = order - 1; j >= 0; j--) {
              i32 accu2 = lsl((i32)b1[j], 16);
              accu2 = add_sat(accu2, mult16x16_shl_sat(parcor[j], b2[j]));
              b1[j + 1] = (int16_t)round16(accu2);
              if (abs32_sat(accu2) == 0x7fffffff) status = 1;
            }
            b1[0] = (int16_t)round16(accu);
            lpc[i] = (int16_t)round16(accu1);
            accu1 = 0;
          }
          if (status) scale_lpc++;
        }
      }
      const int base = win << 7;
      // ixheaacd_calc_max_spectral_line over the filter's region
      int scale_spec;
      {
        i32 mx = 0;
        for (int i = 0; i < tp.size; i++) mx |= abs_nrm(spec[base + tp.start + i]);
        scale_spec = norm32(mx);
      }
      if (tp.apply <= 0) continue;  // the position checks (pns_js_thumb.c:370-386)
      scale_spec = (scale_spec - 4) - scale_lpc;
      // ixheaacd_tns_ar_filter_dec
      int order = tp.order_r;
      if (tp.order & 3) {
        for (int i = tp.order + 1; i <= min(((tp.order & ~3) + 4), 31); i++) lpc[i] = 0;
      }
      int pre = 0;
      int shift_value = scale_lpc;
      if (scale_spec > 0) {
        scale_spec = min(scale_spec, 31);
      } else {
        // reference quirk (pns_js_thumb.c:461): the pre-shift addresses `spec + (win >> 7) + start`, i.e. window 0's region
        pre = min(-scale_spec, 31);
        for (int i = 0; i < tp.size; i++) spec[tp.start + i] = spec[tp.start + i] >> pre;
        scale_spec = 0;
      }
      {
        i32 state[33];
        i32 *sp = spec + base + tp.position;
        const int steps = max(order, tp.size);
        for (int i = 0; i < steps; i++) {
          i32 y = shl32_sat(*sp, scale_spec);
          i32 acc = 0;
          const int jm = i < order ? i : order;
          for (int j = jm; j > 0; j--) {
            acc = add_sat(acc, mul32x16(state[j - 1], (i32)lpc[j]));
            state[j] = state[j - 1];
          }
          y = sub_sat(y, shl32_sat(acc, 1));
          state[0] = shl32_sat(y, shift_value);
          *sp = y >> scale_spec;
          sp += tp.dir;
        }
      }
      if (pre)
        for (int i = 0; i < tp.size; i++) spec[base + tp.start + i] = lsl(spec[base + tp.start + i], pre);
    }
  }
}

cudaError_t launch_aac_spectral(const AacSpectralArgs &args, int num_sms, cudaStream_t stream) {
  long long need = (args.n_units + 7) / 8;
  long long grid = (long long)num_sms * 8;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  aac_stereo_tools_kernel<<<(unsigned)grid, 256, 0, stream>>>(args);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (args.pns_seed) {
    aac_pns_kernel<<<(unsigned)((args.n_units + 127) / 128), 128, 0, stream>>>(args);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  const long long threads = 2 * args.n_units;
  aac_tns_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace xb
