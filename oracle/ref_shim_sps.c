/*
 * oracle/ref_shim_sps.c — TEST INFRASTRUCTURE ONLY.
 *
 * Flat-array entry point around the UNMODIFIED reference's pre-IMDCT spectral stage ixheaacd_channel_pair_process
 * (decoder/ixheaacd_channel.c:602) for AAC-LC elements: the channel-info structs are rebuilt from XAAC_SPS_* records
 * (include/xaac_b200.h), the compiled function runs on the given spectra in place.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#define REF_SHIM_HEADERS_ONLY
#include "ref_headers.h"
#include "ixheaacd_b200_pack_spec.h"

IA_ERRORCODE ixheaacd_channel_pair_process(ia_aac_dec_channel_info_struct *ptr_aac_dec_channel_info[CHANNELS], WORD32 num_ch,
                                           ia_aac_dec_tables_struct *ptr_aac_tables, WORD32 total_channels, WORD32 object_type,
                                           WORD32 aac_spect_data_resil_flag, WORD32 aac_sf_data_resil_flag, WORD32 *in_data,
                                           WORD32 *out_data, void *self_ptr);
extern const ia_aac_dec_block_tables_struct ixheaacd_aac_block_tables;

const void *ref_rom_block_tables(int *bytes) {
  if (bytes) *bytes = 620;
  return &ixheaacd_aac_block_tables;
}

/* spec [n][2][1024] in/out; rec [n][XAAC_SPS_BYTES]; seed [n] in/out (current_seed of the PNS generator); err [n] = the function's
 * return value */
void ref_channel_pair_process_batch(int64_t n, int32_t *spec, const uint8_t *rec, int32_t *seed, int32_t *err) {
  static __thread ia_aac_dec_channel_info_struct ci[2];
  static __thread ia_stereo_info_struct stereo;
  static __thread ia_pns_correlation_info_struct corr;
  static __thread ia_pns_rand_vec_struct rnd;
  static __thread ia_aac_dec_tables_struct tabs;
  static __thread WORD32 scratch[2][1024];
  static __thread WORD16 sfb_idx[4][52], sf[2][128];
  static __thread WORD8 sfb_w[4][52], cb[2][128];
  ia_aac_dec_channel_info_struct *pci[2] = {&ci[0], &ci[1]};
  for (int64_t u = 0; u < n; u++) {
    const uint8_t *r = rec + u * XAAC_SPS_BYTES;
    const int32_t *hdr = (const int32_t *)r;
    const int num_ch = hdr[XAAC_SPS_NUM_CH];
    memset(ci, 0, sizeof(ci));
    memset(&tabs, 0, sizeof(tabs));
    memset(&corr, 0, sizeof(corr));
    memset(&rnd, 0, sizeof(rnd));
    tabs.pstr_block_tables = (ia_aac_dec_block_tables_struct *)&ixheaacd_aac_block_tables;
    memcpy(stereo.ms_used, r + XAAC_SPS_MS_USED, 512);
    memcpy(corr.correlated, r + XAAC_SPS_CORRELATED, 16);
    rnd.current_seed = seed[u];
    for (int c = 0; c < 2; c++) {
      const uint8_t *b = r + XAAC_SPS_CH + c * XAAC_SPS_CH_BYTES;
      const int32_t *w = (const int32_t *)b;
      ia_ics_info_struct *ics = &ci[c].str_ics_info;
      const int ws = w[XAAC_SPS_CH_WINDOW_SEQUENCE];
      ics->window_sequence = (WORD16)ws;
      ics->max_sfb = (WORD16)w[XAAC_SPS_CH_MAX_SFB];
      ics->num_window_groups = (WORD16)w[XAAC_SPS_CH_NUM_WINDOW_GROUPS];
      ics->sampling_rate_index = (WORD16)w[XAAC_SPS_CH_SR_INDEX];
      ics->frame_length = 1024;
      memcpy(ics->window_group_length, b + XAAC_SPS_CH_GROUP_LEN, 8);
      memcpy(cb[c], b + XAAC_SPS_CH_CODE_BOOK, 128);
      memcpy(sf[c], b + XAAC_SPS_CH_SCALE_FACTOR, 256);
      memcpy(&ci[c].str_tns_info, b + XAAC_SPS_CH_TNS, 924);
      memcpy(ci[c].str_pns_info.pns_used, b + XAAC_SPS_CH_PNS_USED, 128);
      ci[c].str_pns_info.pns_active = (UWORD16)w[XAAC_SPS_CH_PNS_ACTIVE];
      ci[c].ptr_code_book = cb[c];
      ci[c].ptr_scale_factor = sf[c];
      ci[c].ptr_spec_coeff = spec + u * 2048 + c * 1024;
      ci[c].pstr_stereo_info = &stereo;
      ci[c].pstr_pns_corr_info = &corr;
      ci[c].pstr_pns_rand_vec_data = &rnd;
      ci[c].scratch_buf_ptr = scratch[c];
      ci[c].common_window = (WORD16)hdr[XAAC_SPS_COMMON_WINDOW];
      if (c < num_ch && ws >= 0 && ws < 4) {
        const int cnt = ws == 2 ? 16 : 52;
        memcpy(sfb_idx[ws], b + XAAC_SPS_CH_SFB_INDEX, (size_t)cnt * 2);
        for (int i = 0; i + 1 < cnt; i++) sfb_w[ws][i] = (WORD8)(sfb_idx[ws][i + 1] - sfb_idx[ws][i]);
        tabs.str_aac_sfb_info[ws].sfb_index = sfb_idx[ws];
        tabs.str_aac_sfb_info[ws].sfb_width = sfb_w[ws];
      }
    }
    memcpy(tabs.sfb_long_table, sfb_idx[0], sizeof(tabs.sfb_long_table));
    err[u] = ixheaacd_channel_pair_process(pci, num_ch, &tabs, num_ch, 2 /* AOT_AAC_LC */, 0, 0, NULL, NULL, NULL);
    seed[u] = rnd.current_seed;
  }
}
