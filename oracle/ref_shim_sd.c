/*
 * oracle/ref_shim_sd.c — TEST INFRASTRUCTURE ONLY.
 *
 * Flat-record entry point around the UNMODIFIED reference's SBR side-info dequantisation ixheaacd_dec_sbrdata
 * (decoder/ixheaacd_env_dec.c:628): the header / frame-data / previous-frame structs are rebuilt from XAAC_SD_* records
 * (include/xaac_b200.h) with the product-side packing header, the compiled function runs, the structs are packed back.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#define REF_SHIM_HEADERS_ONLY
#include "ref_headers.h"
#include "ixheaacd_b200_pack_sd.h"

IA_ERRORCODE ixheaacd_dec_sbrdata(ia_sbr_header_data_struct *ptr_header_data_ch_0, ia_sbr_header_data_struct *ptr_header_data_ch_1,
                                  ia_sbr_frame_info_data_struct *ptr_sbr_data_ch_0,
                                  ia_sbr_prev_frame_data_struct *ptr_prev_data_ch_0,
                                  ia_sbr_frame_info_data_struct *ptr_sbr_data_ch_1,
                                  ia_sbr_prev_frame_data_struct *ptr_prev_data_ch_1, ixheaacd_misc_tables *ptr_common_tables,
                                  WORD32 ldmps_present, WORD32 audio_object_type, WORD32 ec_flag);
extern const ixheaacd_misc_tables ixheaacd_str_fft_n_transcendent_tables;

/* rec [n][XAAC_SD_WORDS] in/out */
void ref_dec_sbrdata_batch(int64_t n, int16_t *rec) {
  static __thread ia_sbr_header_data_struct h[2];
  static __thread ia_freq_band_data_struct fb[2];
  static __thread ia_sbr_frame_info_data_struct f[2];
  static __thread ia_sbr_prev_frame_data_struct p[2];
  for (int64_t u = 0; u < n; u++) {
    int16_t *r = rec + u * XAAC_SD_WORDS;
    const int two = r[XAAC_SD_NUM_CH] == 2, shared = two && r[XAAC_SD_SHARED_HDR];
    memset(h, 0, sizeof(h));
    memset(fb, 0, sizeof(fb));
    memset(p, 0, sizeof(p));
    /* the frame-data structs (100 KB each, mostly PVC / float scratch the function never reads on this path) are zero-initialised
     * thread-locals; every member the function reads is rewritten from the record below, so they are not cleared per record —
     * the bench's CPU arm times this loop */
    for (int c = 0; c < (two ? 2 : 1); c++) {
      const int16_t *b = r + XAAC_SD_CH + c * XAAC_SD_CH_WORDS;
      ia_sbr_header_data_struct *hh = &h[shared ? 0 : c];
      if (!(shared && c == 1)) {
        hh->pstr_freq_band_data = &fb[c];
        fb[c].num_sf_bands[0] = b[XAAC_SDC_NUM_SF_LO];
        fb[c].num_sf_bands[1] = b[XAAC_SDC_NUM_SF_HI];
        fb[c].num_nf_bands = b[XAAC_SDC_NUM_NF];
        hh->num_time_slots = b[XAAC_SDC_NUM_TIME_SLOTS];
        hh->amp_res = b[XAAC_SDC_HDR_AMP_RES];
      }
      b200_sd_unpack_ch(b, hh, &f[c], &p[c]);
      p[c].amp_res = b[XAAC_SDC_PREV_AMP_RES];
      p[c].end_position = b[XAAC_SDC_PREV_END_POS];
      p[c].max_qmf_subband_aac = b[XAAC_SDC_PREV_MAX_QMF];
      p[c].coupling_mode = b[XAAC_SDC_PREV_COUPLING];
      for (int i = 0; i < 10; i++) p[c].sbr_invf_mode[i] = b[XAAC_SDC_PREV_INVF + i];
    }
    if (shared) { /* the flags of the shared header are those of channel block 0 */
      h[0].err_flag = r[XAAC_SD_CH + XAAC_SDC_ERR_FLAG];
      h[0].err_flag_prev = r[XAAC_SD_CH + XAAC_SDC_ERR_FLAG_PREV];
    }
    ia_sbr_header_data_struct *h1 = two ? (shared ? &h[0] : &h[1]) : &h[0];
    IA_ERRORCODE rc = ixheaacd_dec_sbrdata(&h[0], h1, &f[0], &p[0], two ? &f[1] : NULL, two ? &p[1] : NULL,
                                           (ixheaacd_misc_tables *)&ixheaacd_str_fft_n_transcendent_tables, 0, AOT_SBR, 0);
    int16_t keep[8];
    memcpy(keep, r, sizeof(keep));
    {
      int16_t *b0 = r + XAAC_SD_CH, *b1 = b0 + XAAC_SD_CH_WORDS;
      b200_sd_pack_ch(b0, &h[0], &f[0], &p[0]);
      if (two) {
        int16_t e = b1[XAAC_SDC_ERR_FLAG], ep = b1[XAAC_SDC_ERR_FLAG_PREV];
        fb[1] = shared ? fb[0] : fb[1];
        if (shared) { /* keep channel block 1's own (unused) header words as they came in */
          ia_sbr_header_data_struct tmp = h[0];
          int16_t lo = b1[XAAC_SDC_NUM_SF_LO], hi = b1[XAAC_SDC_NUM_SF_HI], nf = b1[XAAC_SDC_NUM_NF], ts = b1[XAAC_SDC_NUM_TIME_SLOTS],
                  ar = b1[XAAC_SDC_HDR_AMP_RES];
          b200_sd_pack_ch(b1, &tmp, &f[1], &p[1]);
          b1[XAAC_SDC_NUM_SF_LO] = lo; b1[XAAC_SDC_NUM_SF_HI] = hi; b1[XAAC_SDC_NUM_NF] = nf; b1[XAAC_SDC_NUM_TIME_SLOTS] = ts;
          b1[XAAC_SDC_HDR_AMP_RES] = ar; b1[XAAC_SDC_ERR_FLAG] = e; b1[XAAC_SDC_ERR_FLAG_PREV] = ep;
        } else {
          b200_sd_pack_ch(b1, &h[1], &f[1], &p[1]);
        }
      }
    }
    memcpy(r, keep, sizeof(keep));
    r[XAAC_SD_ERR] = rc == 0 ? 0 : (rc == (IA_ERRORCODE)-1 ? 2 : 1);
  }
}

/* ixheaacd_decode_ps_data (decoder/ixheaacd_ps_bitdec.c:98) on XAAC_PSD_* records, in place */
VOID ixheaacd_decode_ps_data(ia_ps_dec_struct *ptr_ps_dec, WORD32 frame_size);
void ref_decode_ps_data_batch(int64_t n, int16_t *rec) {
  static __thread ia_ps_dec_struct ps;
  for (int64_t u = 0; u < n; u++) {
    int16_t *r = rec + u * XAAC_PSD_WORDS;
    memset(&ps, 0, sizeof(ps));
    b200_psd_unpack(r, &ps);
    ps.enable_iid = r[XAAC_PSD_ENABLE_IID];
    ps.enable_icc = r[XAAC_PSD_ENABLE_ICC];
    ps.iid_mode = r[XAAC_PSD_IID_MODE];
    ps.icc_mode = r[XAAC_PSD_ICC_MODE];
    ps.iid_quant = r[XAAC_PSD_IID_QUANT];
    ps.frame_class = r[XAAC_PSD_FRAME_CLASS];
    for (int i = 0; i < 5; i++) {
      ps.iid_dt[i] = r[XAAC_PSD_IID_DT + i];
      ps.icc_dt[i] = r[XAAC_PSD_ICC_DT + i];
    }
    const int frame_size = r[XAAC_PSD_FRAME_SIZE];
    ixheaacd_decode_ps_data(&ps, frame_size);
    b200_psd_pack(r, &ps, frame_size);
  }
}
