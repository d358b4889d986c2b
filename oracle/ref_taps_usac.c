/*
 * oracle/ref_taps_usac.c — TEST INFRASTRUCTURE ONLY.
 *
 * Stage tap for the USAC frequency-domain core transform of the UNMODIFIED reference decoder, installed with
 * `ld --wrap=ixheaacd_fd_frm_dec` into oracle/_ref/xaacdec_tap (see oracle/ref_taps.c for the mechanism).
 * record: int32 magic 'UFD1', int32 hdr[8] = {ccfl, window_sequence, window_shape, window_shape_prev, td_frame_prev,
 *         fac_data_present, ec_flag, return value}, int32 coef_in[1024], ov_in[1024], out[1024], ov_out[1024]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include "ixheaac_type_def.h"
#include "ixheaacd_interface.h"
#include "ixheaacd_defines.h"
#include "ixheaacd_aac_rom.h"
#include "ixheaacd_bitbuffer.h"
#include "ixheaacd_tns_usac.h"
#include "ixheaacd_cnst.h"
#include "ixheaacd_acelp_info.h"
#include "ixheaacd_td_mdct.h"
#include "ixheaacd_sbrdecsettings.h"
#include "ixheaacd_info.h"
#include "ixheaacd_sbr_common.h"
#include "ixheaacd_drc_data_struct.h"
#include "ixheaacd_drc_dec.h"
#include "ixheaacd_sbrdecoder.h"
#include "ixheaacd_mps_polyphase.h"
#include "ixheaac_sbr_const.h"
#include "ixheaacd_pulsedata.h"
#include "ixheaacd_pns.h"
#include "ixheaacd_lt_predict.h"
#include "ixheaacd_ec_defines.h"
#include "ixheaacd_ec_struct_def.h"
#include "ixheaacd_main.h"

WORD32 __real_ixheaacd_fd_frm_dec(ia_usac_data_struct *usac_data, WORD32 i_ch);

WORD32 __wrap_ixheaacd_fd_frm_dec(ia_usac_data_struct *ud, WORD32 ch) {
  static FILE *fp = NULL;
  static int tried = 0, count = 0, lim = 0;
  if (!tried) {
    const char *p = getenv("XAAC_TAP_FILE"), *s = getenv("XAAC_TAP_STAGES"), *m = getenv("XAAC_TAP_MAX");
    tried = 1;
    lim = m ? atoi(m) : 1000000;
    if (p && *p && s && strstr(s, "ufd")) {
      static char name[1024];
      snprintf(name, sizeof(name), "%s.ufd", p);
      fp = fopen(name, "wb");
    }
  }
  const int rec = fp && count < lim && ud->ccfl == 1024;
  static int32_t coef_in[1024], ov_in[1024];
  int32_t hdr[9];
  if (rec) {
    hdr[0] = 0x31444655;
    hdr[1] = ud->ccfl;
    hdr[2] = ud->window_sequence[ch];
    hdr[3] = ud->window_shape[ch];
    hdr[4] = ud->window_shape_prev[ch];
    hdr[5] = ud->td_frame_prev[ch];
    hdr[6] = ud->fac_data_present[ch];
    hdr[7] = ud->ec_flag;
    memcpy(coef_in, ud->coef_fix[ch], sizeof(coef_in));
    memcpy(ov_in, ud->overlap_data_ptr[ch], sizeof(ov_in));
  }
  WORD32 ret = __real_ixheaacd_fd_frm_dec(ud, ch);
  if (rec) {
    hdr[8] = ret;
    fwrite(hdr, 4, 9, fp);
    fwrite(coef_in, 4, 1024, fp);
    fwrite(ov_in, 4, 1024, fp);
    fwrite(ud->output_data_ptr[ch], 4, 1024, fp);
    fwrite(ud->overlap_data_ptr[ch], 4, 1024, fp);
    fflush(fp);
    count++;
  }
  return ret;
}
