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

/* ---- fixed-point SBR QMF banks ------------------------------------------------------------------------
 * QMF ROM blob = the leading 3464 bytes of ia_qmf_dec_tables_struct (decoder/ixheaacd_sbr_rom.h:71-95). */
#define XO_QROM_W32 0              /* WORD16[60]  radix-4 twiddles, 32-point */
#define XO_QROM_W16 120            /* WORD16[24]  radix-4 twiddles, 16-point */
#define XO_QROM_DIGREV2_32 168     /* WORD32[4] */
#define XO_QROM_DIGREV4_16 184     /* WORD32[2] */
#define XO_QROM_SINCOS_L64 192     /* WORD16[64] */
#define XO_QROM_ALTSIN_L64 320     /* WORD16[32] */
#define XO_QROM_COSSIN_DS_L32 384  /* WORD16[64] */
#define XO_QROM_SINCOS_L32 512     /* WORD16[32] */
#define XO_QROM_ALTSIN_L32 576     /* WORD16[16] */
#define XO_QROM_TCOSSIN_L32 608    /* WORD16[64] */
#define XO_QROM_POST_FFT 736       /* WORD16[18] */
#define XO_QROM_DCT23_TW 772       /* WORD16[66] */
#define XO_QROM_QMF_C 904          /* WORD16[1280] */
#define XO_QROM_BYTES 3464

void xo_cos_sin_mod(const uint8_t *qrom, int32_t *subband, int no_channels);
void xo_synt_qmffilt_hq(const uint8_t *qrom, int32_t *matrix, int16_t *filter_states, int32_t *drc_offset,
                        int32_t *filter_pos, const int32_t *sf, int lsb, int usb, int split, int16_t *time_out,
                        int ch_fac);
int xo_anal_qmffilt_hq(const uint8_t *qrom, const int16_t *time_in, int ch_fac, int16_t *states, int32_t *pos,
                       int32_t *filter_pos, int usb, int32_t *matrix);
void xo_synt_qmffilt_hq_batch(const uint8_t *qrom, int32_t *matrix, int16_t *filter_states, int32_t *drc_offset,
                              int32_t *filter_pos, const int32_t *sf, const int32_t *lsb, const int32_t *usb,
                              int16_t *time_out, int n);
void xo_anal_qmffilt_hq_batch(const uint8_t *qrom, const int16_t *time_in, int16_t *states, int32_t *pos,
                              int32_t *filter_pos, const int32_t *usb, int32_t *matrix, int n);

/* ---- fixed-point HQ HF generator --------------------------------------------------------------------------
 * Per-unit parameter record (WORD16[80]) — the fields of ia_transposer_settings_struct
 * (decoder/ixheaacd_lpp_tran.h:50-57) and the scalar arguments of ixheaacd_hf_generator (lpp_tran.c:956-962). */
#define XO_HF_NUM_PATCHES 0
#define XO_HF_START_PATCH 1
#define XO_HF_STOP_PATCH 2
#define XO_HF_NUM_COLUMNS 3
#define XO_HF_BW_BORDERS 4      /* [10] */
#define XO_HF_PATCH 14          /* [6][6]: src_start, src_end, guard_start, dst_start, dst_end, num_bands */
#define XO_HF_FACTOR 50         /* time_step */
#define XO_HF_NUM_IF_BANDS 51
#define XO_HF_START_IDX 52      /* border_vec[0] */
#define XO_HF_STOP_IDX 53       /* border_vec[num_env] - num_time_slots */
#define XO_HF_INVF 54           /* [10] sbr_invf_mode */
#define XO_HF_INVF_PREV 64      /* [10] */
#define XO_HF_OV_LB_SCALE 74
#define XO_HF_LB_SCALE 75
#define XO_HF_MAX_QMF_SUBBAND 76
#define XO_HF_PRM_WORDS 80

int xo_hf_generator_hq(const int32_t *lpc, int32_t *matrix, const int16_t *prm, int32_t *bw_prev);
void xo_hf_generator_hq_batch(const int32_t *lpc, int32_t *matrix, const int16_t *prm, int32_t *bw_prev,
                              int32_t *hb_scale, int n);

/* ---- fixed-point HQ envelope adjuster ---------------------------------------------------------------------
 * ROMs: env_rom = the host's ia_env_calc_tables_struct (decoder/ixheaacd_sbr_rom.h:59-68, 2404 bytes);
 *       misc_rom = the leading 2470 bytes of ixheaacd_misc_tables (decoder/ixheaacd_common_rom.h:27-44). */
#define XO_EROM_LIM_GAINS 0   /* WORD16[8]  */
#define XO_EROM_SMOOTH 24     /* WORD16[4]  */
#define XO_EROM_INV_INT 32    /* WORD16[49] */
#define XO_EROM_RAND_PH 132   /* WORD32[512+56] */
#define XO_EROM_BYTES 2404
#define XO_MROM_INV_TABLE 1444  /* WORD16[256] */
#define XO_MROM_SQRT_TABLE 1956 /* WORD16[257] */
#define XO_MROM_BYTES 2470
/* ia_sbr_scale_fact_struct (decoder/ixheaacd_sbr_scale.h:23-31) as WORD16[8] */
#define XO_SF_LB 0
#define XO_SF_ST_LB 1
#define XO_SF_OV_LB 2
#define XO_SF_HB 3
#define XO_SF_OV_HB 4
#define XO_SF_ST_SYN 5
#define XO_SF_PS 6
/* per-frame SBR side-info record (WORD16[XO_ENV_PRM_WORDS]): the fields of ia_sbr_header_data_struct,
 * ia_freq_band_data_struct (decoder/ixheaacd_env_extr_part.h:33-100), ia_frame_info_struct and
 * ia_sbr_frame_info_data_struct (decoder/ixheaacd_env_extr.h:54-120) the fixed-point SBR stage reads */
#define XO_ENV_NUM_TIME_SLOTS 0
#define XO_ENV_TIME_STEP 1
#define XO_ENV_CHANNEL_MODE 2
#define XO_ENV_LIMITER_GAINS 3
#define XO_ENV_INTERPOL_FREQ 4
#define XO_ENV_SMOOTHING_MODE 5
#define XO_ENV_NUM_SF_LO 6
#define XO_ENV_NUM_SF_HI 7
#define XO_ENV_NUM_NF_BANDS 8
#define XO_ENV_SUB_BAND_START 9
#define XO_ENV_SUB_BAND_END 10
#define XO_ENV_NUM_LF_BANDS 11
#define XO_ENV_NUM_ENV 12
#define XO_ENV_TRANSIENT_ENV 13
#define XO_ENV_MAX_QMF_SUBBAND 14
#define XO_ENV_MAX_QMF_SUBBAND_PREV 15
#define XO_ENV_BORDER_VEC 16       /* [9]  */
#define XO_ENV_FREQ_RES 25         /* [8]  */
#define XO_ENV_NOISE_BORDER_VEC 33 /* [3]  */
#define XO_ENV_LIM_TBL 36          /* [13] */
#define XO_ENV_FREQ_LO 49          /* [29] */
#define XO_ENV_FREQ_HI 78          /* [57] */
#define XO_ENV_FREQ_NOISE 135      /* [6]  */
#define XO_ENV_NOISE_FLOOR 141     /* [10] */
#define XO_ENV_ADD_HARMONICS 151   /* [56] */
#define XO_ENV_SF_ARR 207          /* [448] int_env_sf_arr */
#define XO_ENV_PRM_WORDS 656
/* ia_sbr_calc_env_struct (decoder/ixheaacd_env_calc.h:24-33) as WORD16[XO_ENV_ST_WORDS]:
 * filt_buf_me[112], filt_buf_noise_m[56], filt_buf_noise_e, start_up, ph_index, tansient_env_prev, harm_index,
 * harm_flags_prev[56] */
#define XO_ENV_ST_FILT_ME 0
#define XO_ENV_ST_FILT_NOISE 112
#define XO_ENV_ST_NOISE_E 168
#define XO_ENV_ST_START_UP 169
#define XO_ENV_ST_PH_INDEX 170
#define XO_ENV_ST_TRANS_PREV 171
#define XO_ENV_ST_HARM_INDEX 172
#define XO_ENV_ST_HARM_PREV 173
#define XO_ENV_ST_WORDS 232

int xo_expsubbandsamples_hq(const int32_t *matrix, int b0, int b1, int s0, int s1);
void xo_adjust_scale_hq(int32_t *matrix, int b0, int b1, int s0, int s1, int shift);
int xo_calc_sbrenvelope_hq(const uint8_t *env_rom, const uint8_t *misc_rom, const int16_t *prm, int16_t *sf,
                           int16_t *state, int32_t *matrix);
void xo_calc_sbrenvelope_hq_batch(const uint8_t *env_rom, const uint8_t *misc_rom, const int16_t *prm, int16_t *sf,
                                  int16_t *state, int32_t *matrix, int32_t *err, int n);

/* ---- whole fixed-point HQ SBR stage (ixheaacd_sbr_dec) and parametric stereo -------------------------------
 * Per-channel SBR state blob, WORD16[XO_SBR_ST_WORDS] (32-bit members at even offsets, little endian):
 * what ia_sbr_dec_struct carries from frame to frame on this path (SURVEY.md §8a "State per stream"). */
#define XO_SBR_ST_ANAL_STATES 0   /* [320]  str_codec_qmf_bank.anal_filter_states */
#define XO_SBR_ST_ANAL_POS 320    /* [2]    {core_samples_buffer - anal_filter_states, filter_pos - qmf_c} */
#define XO_SBR_ST_SYN_POS 322     /* [2]    {ixheaacd_drc_offset, filter_pos_syn - qmf_c} */
#define XO_SBR_ST_SF 324          /* [8]    str_sbr_scale_fact (XO_SF_*) */
#define XO_SBR_ST_MISC 332        /* [16]   XO_SBR_MISC_* */
#define XO_SBR_ST_ENV 348         /* [232]  str_sbr_calc_env (XO_ENV_ST_*) */
#define XO_SBR_ST_SYN_STATES 580  /* [1280] str_synthesis_qmf_bank.filter_states */
#define XO_SBR_ST_BW_PREV 1860    /* WORD32[6]      str_hf_generator.bw_array_prev */
#define XO_SBR_ST_LPC 1872        /* WORD32[2][128] lpc_filt_states_real[i] | _imag[i] */
#define XO_SBR_ST_OV 2384         /* WORD32[6][128] ptr_sbr_overlap_buf: re[64] | im[64] per overlap slot */
#define XO_SBR_ST_WORDS 3920
#define XO_SBR_MISC_MAX_QMF_PREV 0  /* prev_frame_data.max_qmf_subband_aac */
#define XO_SBR_MISC_END_POS_PREV 1  /* prev_frame_data.end_position */
#define XO_SBR_MISC_INVF_PREV 2     /* [10] prev_frame_data.sbr_invf_mode */
#define XO_SBR_MISC_CODEC_USB 12    /* str_codec_qmf_bank.usb */
#define XO_SBR_MISC_SYN_LSB 13      /* str_synthesis_qmf_bank.lsb */
#define XO_SBR_MISC_SYN_USB 14      /* str_synthesis_qmf_bank.usb */
/* Per-frame side-info record for the whole stage, WORD16[XO_SIDE_WORDS] */
#define XO_SIDE_ENV 0        /* [656] XO_ENV_* (MAX_QMF_SUBBAND_PREV is taken from the state) */
#define XO_SIDE_HF 656       /* [80]  XO_HF_*: transposer settings, NUM_IF_BANDS, INVF; the rest is derived */
#define XO_SIDE_APPLY 736    /* apply_processing (sync_state == SBR_ACTIVE) */
#define XO_SIDE_PS 737       /* 0: no PS; 1: PS active (channel_mode == PS_STEREO), rotation as the reference's own
                                x86-64 gcc build executes it; 2: PS active, rotation as the C source is written
                                (see apply_rot in oracle/src/ps.c) */
#define XO_SIDE_PS_PRM 744   /* [488] XO_PS_PRM_* */
#define XO_SIDE_WORDS 1232
#define XO_PS_PRM_IID_QUANT 0
#define XO_PS_PRM_NUM_ENV 1
#define XO_PS_PRM_BORDER 2   /* [7]     border_position */
#define XO_PS_PRM_IID 9      /* [7][34] iid_par_table */
#define XO_PS_PRM_ICC 247    /* [7][34] icc_par_table */
#define XO_PS_PRM_WORDS 488
/* PS state blob, WORD16[XO_PS_ST_WORDS]: ia_ps_dec_struct delay lines, mixing matrices, transient detector,
 * hybrid delay lines (decoder/ixheaacd_ps_dec.h:97-141) + the right-channel synthesis bank */
#define XO_PS_ST_AP 0             /* [2][64]    delay_buf_qmf_ap_re_im */
#define XO_PS_ST_LD 128           /* [14][24]   delay_buf_qmf_ld_re_im */
#define XO_PS_ST_SD 464           /* [64]       delay_buf_qmf_sd_re_im */
#define XO_PS_ST_SER 528          /* [5][3][64] delay_buf_qmf_ser_re_im */
#define XO_PS_ST_SUB 1488         /* [2][32]    delay_buf_qmf_sub_re_im */
#define XO_PS_ST_SUB_SER 1552     /* [5][3][32] delay_buf_qmf_sub_ser_re_im */
#define XO_PS_ST_HVEC 2032        /* [6][48]    h11_h12_vec, h21_h22_vec, H11_H12, H21_H22, delta_h11_h12, delta_h21_h22 */
#define XO_PS_ST_IDX 2320         /* [12]       XO_PS_IDX_* */
#define XO_PS_ST_PEAK 2332        /* WORD32[3][20] peak_decay_diff, energy_prev, peak_decay_diff_prev */
#define XO_PS_ST_HYB 2452         /* WORD32[3][2][12] str_hybrid.ptr_qmf_buf_re[b] | _im[b] */
#define XO_PS_ST_SYN_STATES_R 2596 /* [1280]    right synthesis bank filter_states */
#define XO_PS_ST_SYN_POS_R 3876   /* [2] */
#define XO_PS_ST_SF_R 3878        /* [8]        right channel's str_sbr_scale_fact */
#define XO_PS_ST_WORDS 3888
#define XO_PS_IDX_SER 0           /* [3] delay_buf_idx_ser */
#define XO_PS_IDX_DELAY 3
#define XO_PS_IDX_DELAY_LONG 4
#define XO_PS_IDX_SCALE 5         /* delay_buffer_scale */
#define XO_PS_IDX_USB 6
#define XO_PS_IDX_LSB_R 8         /* right synthesis bank lsb / usb */
#define XO_PS_IDX_USB_R 9
/* PS ROM = leading 1230 bytes of ia_ps_tables_struct (decoder/ixheaacd_sbr_rom.h:177-203); WORD16 offsets */
#define XO_PSROM_DECAY_SF 0
#define XO_PSROM_HYB_RESOL 72
#define XO_PSROM_REV_DECAY 75
#define XO_PSROM_REV_DELAY 78
#define XO_PSROM_BORDERS_GROUP 81
#define XO_PSROM_GROUP_SHIFT 104
#define XO_PSROM_GROUP_TO_BIN 110
#define XO_PSROM_HYB_TO_BIN 132
#define XO_PSROM_DELAY_TO_BIN 142
#define XO_PSROM_FRAC_QMF 174
#define XO_PSROM_FRAC_SUB 222
#define XO_PSROM_FRAC_QMF_SER 254
#define XO_PSROM_FRAC_SUB_SER 446
#define XO_PSROM_SCALE 542
#define XO_PSROM_SCALE_FINE 557
#define XO_PSROM_ALPHA 588
#define XO_PSROM_P2_6 596
#define XO_PSROM_P8_13 602
#define XO_PSROM_BYTES 1230

void xo_ps_apply_frame(const uint8_t *env_rom, const uint8_t *misc_rom, const uint8_t *ps_rom, const int16_t *ps_prm,
                       int16_t *ps, int32_t *m, int32_t *right, int usb, int shiftdelay_late, int common_shift,
                       int as_built);
void xo_ps_synth_pair(const uint8_t *qrom, const uint8_t *env_rom, const uint8_t *misc_rom, const uint8_t *ps_rom,
                      const int16_t *ps_prm, int16_t *st, int16_t *ps, int32_t *m, int16_t *out_l, int16_t *out_r,
                      int ch_out, int as_built);
int xo_sbr_dec_hq(const uint8_t *qrom, const uint8_t *env_rom, const uint8_t *misc_rom, const uint8_t *ps_rom,
                  const int16_t *side, int16_t *st, int16_t *ps_st, const int16_t *time_in, int ch_in,
                  int16_t *time_out, int16_t *time_out_r, int ch_out, int32_t *scratch);
void xo_imdct_out_to_pcm16(const int32_t *in, const int8_t *qshift_adj, int16_t *out, int n_units, int mode);

/* ---- low-power (real-valued) SBR path: matrices are [slot][64] real ------------------------------------------ */
void xo_dct3_32(const uint8_t *qrom, int32_t *in, int32_t *out);
int xo_anal_qmffilt_lp(const uint8_t *qrom, const int16_t *time_in, int ch_fac, int16_t *states, int32_t *pos,
                       int32_t *filter_pos, int32_t *matrix);
void xo_synt_qmffilt_lp(const uint8_t *qrom, int32_t *matrix, int16_t *filter_states, int32_t *drc_offset,
                        int32_t *filter_pos, const int32_t *sf, int lsb, int usb, int split, int16_t *time_out,
                        int ch_fac);
int xo_expsubbandsamples_lp(const int32_t *matrix, int b0, int b1, int s0, int s1);
void xo_adjust_scale_lp(int32_t *matrix, int b0, int b1, int s0, int s1, int shift);
int xo_calc_sbrenvelope_lp(const uint8_t *env_rom, const uint8_t *misc_rom, const int16_t *prm, int16_t *sf,
                           int16_t *state, int32_t *matrix, const int16_t *degree_alias);
void xo_low_pow_hf_generator(const int32_t *lpc, int32_t *scratch, const int16_t *prm, int32_t *bw_prev,
                             int16_t *degree_alias, int norm_max);
int xo_sbr_dec_lp(const uint8_t *qrom, const uint8_t *env_rom, const uint8_t *misc_rom, const int16_t *side,
                  int16_t *st, const int16_t *time_in, int ch_in, int16_t *time_out, int ch_out, int32_t *scratch);

/* ---- USAC frequency-domain core transform (usac_fd.c) ------------------------------------------------------------
 * ROM blob (XO_UROM_BYTES): the reference's const tables of the path, concatenated by ref_rom_usac_tables(). */
#define XO_UROM_FFT_TW 0          /* WORD32[514]  ixheaacd_twiddle_table_fft_32x32 */
#define XO_UROM_COS512 2056       /* WORD32[512]  ixheaacd_pre_post_twid_cos_512 */
#define XO_UROM_SIN512 4104       /* WORD32[512]  ixheaacd_pre_post_twid_sin_512 */
#define XO_UROM_COS64 6152        /* WORD32[64]   ixheaacd_pre_post_twid_cos_64 */
#define XO_UROM_SIN64 6408        /* WORD32[64]   ixheaacd_pre_post_twid_sin_64 */
#define XO_UROM_SINE1024 6664     /* WORD32[1024] ixheaacd_sine_win_1024 */
#define XO_UROM_KBD1024 10760     /* WORD32[1024] ixheaacd_kbd_win1024 */
#define XO_UROM_SINE128 14856     /* WORD32[128]  ixheaacd_sine_win_128 */
#define XO_UROM_KBD128 15368      /* WORD32[128]  ixheaacd_kbd_win128 */
#define XO_UROM_BYTES 15880
int xo_usac_complex_fft(const uint8_t *urom, int32_t *xr, int32_t *xi, int npoints, int preshift);
int xo_usac_fd_frm_dec(const uint8_t *urom, int32_t *coef, int32_t *ov, int win_seq, int win_shape, int win_shape_prev,
                       int32_t *out);
void xo_usac_fd_frm_dec_batch(const uint8_t *urom, int32_t *coef, int32_t *ov, const int32_t *win_seq,
                              const int32_t *win_shape, const int32_t *win_shape_prev, int32_t *out, int32_t *err, int n);


/* ---- AAC-LC output stage: peak limiter (peaklim.c) -----------------------------------------------------------------
 * Per-stream state record = ia_peak_limiter_struct (decoder/ixheaacd_peak_limiter_struct_def.h:29-48) as 32-bit words */
#define XO_PL_ATTACK_CONST 0   /* float  attack_constant */
#define XO_PL_RELEASE_CONST 1  /* float  release_constant */
#define XO_PL_GAIN_MOD 2       /* float  gain_modified */
#define XO_PL_MIN_GAIN 3       /* float  min_gain (out) */
#define XO_PL_PSG 4            /* double pre_smoothed_gain (2 words) */
#define XO_PL_ATTACK 6         /* attack_time_samples (<= XO_PL_MAX_ATTACK) */
#define XO_PL_DELAY_IDX 7      /* delayed_input_index */
#define XO_PL_MAX_IDX 8        /* max_idx */
#define XO_PL_CIR 9            /* cir_buf_pnt */
#define XO_PL_LIMITER_ON 10
#define XO_PL_NUM_CH 11        /* 1 or 2 */
#define XO_PL_MAX_BUF 12       /* float[XO_PL_MAX_ATTACK] max_buf */
#define XO_PL_MAX_ATTACK 512
#define XO_PL_DELAYED 524      /* float[2 * XO_PL_MAX_ATTACK] delayed_input, [index][channel] */
#define XO_PL_WORDS 1548
int xo_peak_limiter(int32_t *st, int32_t *samples, int frame_len, const int8_t *qshift_adj, int16_t *pcm16);
void xo_peak_limiter_batch(int32_t *st, int32_t *samples, const int8_t *qshift_adj, int16_t *pcm16, int32_t *err, int ch, int n);


/* ---- eSBR 64-band synthesis bank (esbr_qmf.c) ------------------------------------------------------------------------
 * ROM blob (XO_EROM2_BYTES): members of ia_qmf_dec_tables_struct concatenated by ref_rom_esbr_tables() */
#define XO_EROM2_QMF_C 0          /* WORD32[1280] esbr_qmf_c */
#define XO_EROM2_W32 5120         /* WORD32[60]   esbr_w_32 */
#define XO_EROM2_SINCOS_L64 5360  /* WORD32[64]   esbr_sin_cos_twiddle_l64 */
#define XO_EROM2_ALTSIN_L64 5616  /* WORD32[32]   esbr_alt_sin_twiddle_l64 */
#define XO_EROM2_W16 5744         /* WORD32[24]   esbr_w_16 */
#define XO_EROM2_SINCOS_L32 5840  /* WORD32[32]   esbr_sin_cos_twiddle_l32 */
#define XO_EROM2_ALTSIN_L32 5968  /* WORD32[16]   esbr_alt_sin_twiddle_l32 */
#define XO_EROM2_TCOS_L32 6032    /* WORD32[64]   esbr_t_cos_sin_l32 */
#define XO_EROM2_BYTES 6288
void xo_esbr_synth64(const uint8_t *erom, const float *qmf, int32_t *fs, int32_t *off_io, int32_t *fpos_io, float *out);
void xo_esbr_synth64_batch(const uint8_t *erom, const float *qmf, int32_t *fs, int32_t *pos, float *out, int n);
void xo_esbr_anal32(const uint8_t *erom, const float *time_in, int32_t *states, int32_t *pos_io, int32_t *fpos_io, float *qmf);
void xo_esbr_anal32_batch(const uint8_t *erom, const float *time_in, int32_t *states, int32_t *pos, float *qmf, int n);

/* eSBR hand-overs, esbr_qmf.c */
void xo_esbr_core_to_float(const int32_t *core, float *out, int n);                  /* ixheaacd_ext_ch_ele.c:1040-1046 */
void xo_esbr_pcm16_to_float(const int16_t *pcm, int ch_fac, int ch, float *out, int n); /* ixheaacd_api.c:3384-3437 */
void xo_samples_sat16(const float *in, int ch_fac, int ch, int16_t *pcm, int n);      /* ixheaacd_decode_main.c:82-104 */

/* ---- eSBR float HF generator (esbr_hfgen.c): ixheaacd_generate_hf ------------------------------------------------------
 * QMF buffers are [XO_EHF_ROWS][64] floats; row r is row r - 2 of the pointers the reference passes
 * (qmf_buf_real + SBR_HF_ADJ_OFFSET etc.), i.e. the reference's own arrays from their first row. par[] words: */
#define XO_EHF_ROWS 40
#define XO_EHF_NUM_MF 0        /* pstr_freq_band_data->num_mf_bands */
#define XO_EHF_NUM_IF 1        /* pstr_freq_band_data->num_nf_bands */
#define XO_EHF_SB_START 2      /* pstr_freq_band_data->sub_band_start */
#define XO_EHF_BORDER_FIRST 3  /* str_frame_info_details.border_vec[0] */
#define XO_EHF_BORDER_LAST 4   /* border_vec[num_env] */
#define XO_EHF_HBE_FLAG 5
#define XO_EHF_PATCHING_MODE 6 /* sbr_patching_mode */
#define XO_EHF_FS 7            /* out_sampling_freq */
#define XO_EHF_PRE_PROC 8      /* pre_proc_flag: ixheaacd_pre_processing (libm log10 / pow in double) */
#define XO_EHF_USF4 9          /* is_usf_4: refused (76-row covariance) */
#define XO_EHF_MPS_SBR 10      /* mps_sbr_flag */
#define XO_EHF_COV_COUNT 11
#define XO_EHF_INVF 16         /* sbr_invf_mode[5] */
#define XO_EHF_INVF_PREV 21    /* sbr_invf_mode_prev[5] */
#define XO_EHF_INVF_TBL 26     /* freq_band_tbl_noise[1..5] */
#define XO_EHF_FMASTER 32      /* f_master_tbl[57] */
#define XO_EHF_PAR_WORDS 96
/* returns 0 / -1 like the reference, -2 for a configuration outside the restated subset.  pv_* may be NULL (no HBE).
 * patch_out[0] = num_patches, [1..7] = start_subband[]; bw_prev[6] in/out (bw_array_prev) */
int xo_esbr_generate_hf(const float *src_re, const float *src_im, const float *pv_re, const float *pv_im, float *dst_re,
                        float *dst_im, const int32_t *par, float *bw_prev, int32_t *patch_out);
void xo_esbr_generate_hf_batch(const float *src_re, const float *src_im, const float *pv_re, const float *pv_im,
                               float *dst_re, float *dst_im, const int32_t *par, float *bw_prev, int32_t *patch_out,
                               int32_t *err, int n);

/* ---- eSBR float envelope adjuster (esbr_envcalc.c): ixheaacd_sbr_env_calc, ORIG_SBR branch ----------------------------
 * ipar[] words (in/out where noted), fpar[] floats, state[] floats; QMF rows as for the HF generator (row r = row r - 2). */
#define XO_EEC_SB_START 0
#define XO_EEC_SB_END 1
#define XO_EEC_NUM_ENV 2
#define XO_EEC_TRANS_ENV 3
#define XO_EEC_SHORT_PREV 4   /* env_short_flag_prev, in/out */
#define XO_EEC_NUM_NOISE_ENV 5
#define XO_EEC_NUM_SF_LO 6
#define XO_EEC_NUM_SF_HI 7
#define XO_EEC_NUM_NF 8
#define XO_EEC_SMOOTHING_MODE 9
#define XO_EEC_INTERPOL_FREQ 10
#define XO_EEC_LIMITER_BANDS 11
#define XO_EEC_LIMITER_GAINS 12
#define XO_EEC_HARM_INDEX 13  /* in/out */
#define XO_EEC_PHASE_INDEX 14 /* in/out */
#define XO_EEC_START_UP 15    /* pstr_sbr_header->esbr_start_up, in/out */
#define XO_EEC_RESET 16       /* reset_flag (needs XO_EEC_LIM_REBUILT: the limiter tables are rebuilt by the host) */
#define XO_EEC_SBR_MODE 17    /* must be ORIG_SBR (1) */
#define XO_EEC_USF4 18        /* must be 0 */
#define XO_EEC_PATCHING_CHANGED 19 /* sbr_patching_mode != prev_sbr_patching_mode */
#define XO_EEC_LIM_REBUILT 20 /* the host has rebuilt lim_table / gate_mode for this reset / patching-change frame */
#define XO_EEC_BORDER 24      /* border_vec[9] */
#define XO_EEC_FREQ_RES 33    /* freq_res[8] */
#define XO_EEC_NOISE_BORDER 41 /* noise_border_vec[3] */
#define XO_EEC_INTER_TES 44   /* inter_temp_shape_mode[8]: must be 0 (gamma = 0) */
#define XO_EEC_GATE_MODE 52   /* gate_mode[4] */
#define XO_EEC_LIM_TABLE 56   /* lim_table[4][13] */
#define XO_EEC_TBL_NOISE 108  /* freq_band_tbl_noise[6] */
#define XO_EEC_TBL_LO 116     /* freq_band_tbl_lo[29] */
#define XO_EEC_TBL_HI 148     /* freq_band_tbl_hi[57] */
#define XO_EEC_ADD_HARM 208   /* add_harmonics[56] */
#define XO_EEC_HARM_PREV 264  /* harm_flag_prev[64] as bytes (16 words), in/out */
#define XO_EEC_IPAR_WORDS 288
#define XO_EEC_SFB_NRG 0      /* flt_env_sf_arr[448] */
#define XO_EEC_NOISE_FLOOR 448 /* flt_noise_floor[10] */
#define XO_EEC_FPAR_WORDS 464
#define XO_EEC_STATE_WORDS 640 /* e_gain[5][64] | noise_buf[5][64], in/out */
#define XO_EEC_NUM_ROWS_MAX 38     /* slots 0..37 of the 40-row buffers */
#define XO_EEC_RPHASE_WORDS 1024 /* ROM: ixheaac_random_phase[512][2] */
int xo_esbr_env_calc(const float *rphase, float *re, float *im, int32_t *ipar, const float *fpar, float *state);
void xo_esbr_env_calc_batch(const float *rphase, float *re, float *im, int32_t *ipar, const float *fpar, float *state,
                            int32_t *err, int n);

/* ---- eSBR QMF harmonic transposer (esbr_hbe.c): ixheaacd_qmf_hbe_apply (decoder/ixheaacd_hbe_trans.c:224-296) --------
 * 2:1 system, 32 QMF columns per call (no_bins = 32, qmf_voc_columns = 16), synth_size 4, 8, 12, 16 (FFT banks) and 20
 * (direct-form banks; for this size the reference's FFT pointers stay NULL, so it re-initialises the instance on every
 * call, hbe_trans.c:240-248, 162-167: synth_buf / analy_buf start from zero each frame — restated as such).  ROM blob = the reference's global float tables concatenated by ref_rom_hbe_tables(), word offsets: */
#define XO_HROM_WIN 0          /* ixheaac_sub_samp_qmf_window_coeff[1560] (common/ixheaac_esbr_rom.c) */
#define XO_HROM_SYNCOS 1560    /* ixheaac_synth_cos_table_kl_4[16] | _8[32] | _12[48] | _16[64] */
#define XO_HROM_ANACS 1720     /* ixheaac_analy_cos_sin_table_kl_8[32] | _16[64] | _24[96] | _32[128] */
#define XO_HROM_COSTRANS 2040  /* ixheaac_cos_table_trans_qmf[7][64] */
#define XO_HROM_FFTTW 2488     /* ixheaac_twiddle_table_fft_float[514] (+2 pad) */
#define XO_HROM_TW24 3004      /* ixheaac_twidle_tbl_24[32] */
#define XO_HROM_TW48 3036      /* ixheaac_twidle_tbl_48[64] */
#define XO_HROM_PVCOS 3100     /* ixheaac_phase_vocoder_cos_table[64] */
#define XO_HROM_PVSIN 3164     /* ixheaac_phase_vocoder_sin_table[64] */
#define XO_HROM_INTERP 3228    /* ixheaac_hbe_post_anal_proc_interp_coeff[4][2] */
#define XO_HROM_SELCASE 3236   /* ixheaac_sel_case[5][8] */
#define XO_HROM_XP2 3276       /* ixheaac_hbe_x_prod_cos_table_trans_2[512] */
#define XO_HROM_XP3 3788       /* ..._trans_3[512] */
#define XO_HROM_XP4 4300       /* ..._trans_4[512] */
#define XO_HROM_XP41 4812      /* ..._trans_4_1[512] */
#define XO_HROM_SYN20 5324     /* ixheaac_synth_cos_table_kl_20[800] (direct-form bank, synth_size 20) */
#define XO_HROM_ANA40 6124     /* ixheaac_analy_cos_sin_table_kl_40[3200] */
#define XO_HROM_WORDS 9324
/* cfg[] words (ia_esbr_hbe_txposer_struct members set by ixheaacd_qmf_hbe_data_reinit, + the call's pitch_in_bins) */
#define XO_HBE_SYNTH_SIZE 0
#define XO_HBE_K_START 1
#define XO_HBE_START_BAND 2
#define XO_HBE_END_BAND 3
#define XO_HBE_MAX_STRETCH 4
#define XO_HBE_PITCH 5
#define XO_HBE_USF4 6          /* upsamp_4_flag: must be 0 */
#define XO_HBE_REINIT 7        /* informational (taps): the instance's FFT pointers were NULL at entry, i.e. the reference
                                * re-initialised it inside the call (always the case for synth_size 20) */
#define XO_HBE_XOVER 8         /* x_over_qmf[6] */
#define XO_HBE_CFG_WORDS 16
/* state[] floats, in/out: what the transposer carries from one call to the next */
#define XO_HBE_ST_TAIL 0       /* [32]  ptr_input_buf[no_bins * synth_size ..+ synth_size) (the next call's first samples) */
#define XO_HBE_ST_SYNTH 32     /* [384] synth_buf[0..18 * synth_size) */
#define XO_HBE_ST_ANAL 416     /* [384] analy_buf[0..18 * synth_size) */
#define XO_HBE_ST_QIN 800      /* [12][128] qmf_in_buf rows 16..27 (the next call's rows 0..11) */
#define XO_HBE_ST_QOUT 2336    /* [10][128] qmf_out_buf rows 32..41 (the next call's rows 0..9; rows above are zero) */
#define XO_HBE_ST_WORDS 3616
/* qmf_re / qmf_im: [32][64] the frame's new QMF slots; pv_re / pv_im: [32][64], bands start_band..end_band-1 written.
 * Returns 0, the reference's error codes, or -2 (outside the restated subset). */
int xo_esbr_hbe_apply(const float *rom, const int32_t *cfg, float *state, const float *qmf_re, const float *qmf_im,
                      float *pv_re, float *pv_im);
void xo_esbr_hbe_apply_batch(const float *rom, const int32_t *cfg, float *state, const float *qmf_re, const float *qmf_im,
                             float *pv_re, float *pv_im, int32_t *err, int n);

#endif
