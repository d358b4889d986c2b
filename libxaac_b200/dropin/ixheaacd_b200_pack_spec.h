/*
 * libxaac_b200/dropin/ixheaacd_b200_pack_spec.h — reference-side half of the pre-IMDCT spectral stage hand-over
 * (ixheaacd_channel_pair_process, decoder/ixheaacd_channel.c:602-718): ia_aac_dec_channel_info_struct[] -> XAAC_SPS_* record.
 * The record carries the reference's own plain-data members byte for byte (ms_used, code books, scale factors, the TNS info
 * struct, the sfb offsets of the element's window sequence).  Used by ixheaacd_b200_glue.c and by oracle/ref_shim_sps.c.
 */
#ifndef IXHEAACD_B200_PACK_SPEC_H
#define IXHEAACD_B200_PACK_SPEC_H
#include <string.h>
#include "ixheaacd_b200_ref_headers.h"
#include "xaac_b200.h"

/* 0, or -1 when the element is outside what the device stage covers (the caller then runs the reference's own code) */
static int b200_sps_pack(uint8_t *rec, ia_aac_dec_channel_info_struct *ci[], int num_ch, ia_aac_dec_tables_struct *t) {
  int32_t *hdr = (int32_t *)rec;
  if (num_ch < 1 || num_ch > 2 || sizeof(ia_tns_info_aac_struct) != 924) return -1;
  memset(rec, 0, XAAC_SPS_BYTES);
  hdr[XAAC_SPS_NUM_CH] = num_ch;
  hdr[XAAC_SPS_COMMON_WINDOW] = ci[0]->common_window;
  if (ci[0]->pstr_stereo_info) memcpy(rec + XAAC_SPS_MS_USED, ci[0]->pstr_stereo_info->ms_used, 512);
  if (ci[0]->pstr_pns_corr_info) memcpy(rec + XAAC_SPS_CORRELATED, ci[0]->pstr_pns_corr_info->correlated, 16);
  for (int c = 0; c < num_ch; c++) {
    uint8_t *b = rec + XAAC_SPS_CH + c * XAAC_SPS_CH_BYTES;
    int32_t *w = (int32_t *)b;
    const ia_ics_info_struct *ics = &ci[c]->str_ics_info;
    const int ws = ics->window_sequence;
    if (ics->frame_length != 1024 || ws < 0 || ws > 3) return -1;
    w[XAAC_SPS_CH_WINDOW_SEQUENCE] = ws;
    w[XAAC_SPS_CH_MAX_SFB] = ics->max_sfb;
    w[XAAC_SPS_CH_NUM_WINDOW_GROUPS] = ics->num_window_groups;
    w[XAAC_SPS_CH_PNS_ACTIVE] = ci[c]->str_pns_info.pns_active;
    if (ics->sampling_rate_index < 0 || ics->sampling_rate_index > 11) return -1;
    w[XAAC_SPS_CH_TNS_MAX_BANDS] = t->pstr_block_tables->tns_max_bands_tbl[ics->sampling_rate_index][ws == 2];
    w[XAAC_SPS_CH_SR_INDEX] = ics->sampling_rate_index;
    memcpy(b + XAAC_SPS_CH_GROUP_LEN, ics->window_group_length, 8);
    memcpy(b + XAAC_SPS_CH_CODE_BOOK, ci[c]->ptr_code_book, 128);
    memcpy(b + XAAC_SPS_CH_SCALE_FACTOR, ci[c]->ptr_scale_factor, 256);
    memcpy(b + XAAC_SPS_CH_TNS, &ci[c]->str_tns_info, 924);
    memcpy(b + XAAC_SPS_CH_PNS_USED, ci[c]->str_pns_info.pns_used, 128);
    {
      const WORD16 *idx = t->str_aac_sfb_info[ws].sfb_index;
      const int n = ws == 2 ? 16 : 52; /* sfb_short_table[16] / sfb_long_table[52] (decoder/ixheaacd_aac_rom.h:184-185) */
      memcpy(b + XAAC_SPS_CH_SFB_INDEX, idx, (size_t)n * 2);
    }
  }
  return 0;
}
#endif
