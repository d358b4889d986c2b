/*
 * xaac_b200.h — C-ABI of libxaac_b200.so: B200-native (sm_100a) kernels for the decode-side DSP hot path
 * of ittiam-systems/libxaac.  Plain C, plain pointers and sizes; no CUDA or torch types in the signatures
 * (streams travel as void*).  Host code that stays C — the reference's own bitstream parser — binds to
 * exactly these entry points (see INTEGRATION.md for the function-selector / --wrap stubs).
 *
 * Conventions (mirroring the reference, decoder/ixheaacd_error_standards.h:24-26):
 *   return 0 on success; a value with bit 31 set (XAAC_B200_FATAL) is fatal (CUDA failure, bad argument).
 *   The library never falls back to a CPU path: without a CUDA device every call fails loudly.
 *   The caller owns every buffer.  "_dev" entry points take device pointers and are asynchronous on the
 *   given stream; "_host" entry points take host pointers, copy in, run, copy out and synchronise.
 *
 * Unit of work for the IMDCT stage: one frame x one core channel ("unit"), 1024 spectral lines.
 */
#ifndef XAAC_B200_H
#define XAAC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XAAC_B200_OK 0
#define XAAC_B200_FATAL ((int32_t)0x80000000)
#define XAAC_B200_ERR_CUDA ((int32_t)0x80000001)
#define XAAC_B200_ERR_ARG ((int32_t)0x80000002)
#define XAAC_B200_ERR_NO_ROM ((int32_t)0x80000003)

/* window_sequence codes, decoder/ixheaacd_cnst.h:100-103 */
#define XAAC_ONLY_LONG_SEQUENCE 0
#define XAAC_LONG_START_SEQUENCE 1
#define XAAC_EIGHT_SHORT_SEQUENCE 2
#define XAAC_LONG_STOP_SEQUENCE 3

typedef struct xaac_b200_ctx xaac_b200_ctx;

/* ---- context --------------------------------------------------------------------------------------- */
int32_t xaac_b200_create(xaac_b200_ctx **ctx, int32_t device);
void xaac_b200_destroy(xaac_b200_ctx *ctx);
const char *xaac_b200_last_error(const xaac_b200_ctx *ctx);
int32_t xaac_b200_num_sms(const xaac_b200_ctx *ctx);
/* number of kernel launches issued by this context so far (bench.py's gpu_launches evidence) */
int64_t xaac_b200_launch_count(const xaac_b200_ctx *ctx);
int32_t xaac_b200_sync(xaac_b200_ctx *ctx);
/* Profiling hook (the counterpart of the reference testbench's ARM_PROFILE_HW MCPS printout,
 * test/decoder/ixheaacd_main.c:2198-2229): when enabled, every kernel launch is bracketed by CUDA events on its own
 * stream.  xaac_b200_kernel_times synchronises and writes "kernel:total_ms:launches;" records into buf.
 * Enabling / disabling clears the records. */
int32_t xaac_b200_kernel_timing(xaac_b200_ctx *ctx, int32_t enable);
int32_t xaac_b200_kernel_times(xaac_b200_ctx *ctx, char *buf, size_t buf_bytes);

/* ---- ROM tables ------------------------------------------------------------------------------------
 * The reference hands its const tables to every hot function as pointer arguments
 * (ia_aac_dec_imdct_tables_struct *, decoder/ixheaacd_aac_rom.h:112-168).  The drop-in does
 * the same once per context: `tables` points at the host's ia_aac_dec_imdct_tables_struct (only the leading
 * XAAC_B200_IMDCT_ROM_BYTES are read: cosine_array_2048_256, dig_rev_table8_long/short, fft_twiddle,
 * only_long_window_sine/kbd, only_short_window_sine/kbd). */
#define XAAC_B200_IMDCT_ROM_BYTES 7500
int32_t xaac_b200_set_imdct_rom(xaac_b200_ctx *ctx, const void *tables, size_t bytes);

/* ---- AAC IMDCT + window/overlap-add (batched) -------------------------------------------------------
 * Replaces ixheaacd_imdct_process (decoder/ixheaacd_lpfuncs.c:347-802) for 1024-sample frames, i.e. the
 * selector leaves ixheaacd_calc_max_spectral_line / pretwiddle_compute / imdct_using_fft / post_twiddle /
 * post_twid_overlap_add / over_lap_add1 / over_lap_add2 / spec_to_overlapbuf / overlap_buf_out /
 * overlap_out_copy / neg_shift_spec (decoder/ixheaacd_function_selector.h:61-223) fused into one kernel.
 *
 *   spec        [n_units][1024] WORD32 spectral coefficients (ptr_spec_coeff); not modified
 *   overlap     [n_units][512]  WORD32 ia_aac_dec_overlap_info.ptr_overlap_buf, updated in place
 *   wstate      [n_units][2]    {window_shape, window_sequence} saved from the previous frame
 *                               (ia_aac_dec_overlap_info), updated in place
 *   ics         [n_units][2]    {window_sequence, window_shape} of this frame (ia_ics_info_struct)
 *   out         WORD32 time samples. ch_fac == 1: [n_units][1024]. ch_fac > 1: units u, u+1, .. u+ch_fac-1
 *               are the channels of one frame, written interleaved with stride ch_fac like the reference's
 *               time buffer.
 *   qshift_adj  [n_units] ia_ics_info_struct.qshift_adj as set by the stage (2, 1)
 */
int32_t xaac_b200_imdct_process_dev(xaac_b200_ctx *ctx, const int32_t *d_spec, int32_t *d_overlap,
                                    uint8_t *d_wstate, const uint8_t *d_ics, int32_t *d_out,
                                    int8_t *d_qshift_adj, int64_t n_units, int32_t ch_fac, void *stream);

/* Device-resident per-unit state of the stage (what ia_aac_dec_overlap_info carries from frame to frame:
 * ptr_overlap_buf[512], window_shape, window_sequence; decoder/ixheaacd_channelinfo.h:88-93).  A stream's state
 * stays in HBM between frames; upload/download are the checkpoint/resume path. A new state is the decoder's
 * reset state: zero overlap, sine window, ONLY_LONG_SEQUENCE. */
typedef struct xaac_b200_imdct_state xaac_b200_imdct_state;
int32_t xaac_b200_imdct_state_create(xaac_b200_ctx *ctx, int64_t n_units, xaac_b200_imdct_state **state);
void xaac_b200_imdct_state_destroy(xaac_b200_ctx *ctx, xaac_b200_imdct_state *state);
int32_t xaac_b200_imdct_state_upload(xaac_b200_ctx *ctx, xaac_b200_imdct_state *state, const int32_t *overlap,
                                     const uint8_t *wstate);
int32_t xaac_b200_imdct_state_download(xaac_b200_ctx *ctx, xaac_b200_imdct_state *state, int32_t *overlap,
                                       uint8_t *wstate);

/* Host-buffer entry point: one frame for every unit of `state`.  spec/ics are read from host memory, out and
 * qshift_adj are written to host memory; the batch is chunked and copies overlap with the kernel on internal
 * streams (pinned host memory recommended).  Returns after everything has completed. */
int32_t xaac_b200_imdct_process_host(xaac_b200_ctx *ctx, xaac_b200_imdct_state *state, const int32_t *spec,
                                     const uint8_t *ics, int32_t *out, int8_t *qshift_adj, int32_t ch_fac);

/* ---- fixed-point SBR QMF banks ----------------------------------------------------------------------
 * ROM: `tables` points at the host's ia_qmf_dec_tables_struct (decoder/ixheaacd_sbr_rom.h:71-123), i.e. what the
 * reference passes around as sbr_tables_ptr->qmf_dec_tables_ptr; only the leading XAAC_B200_QMF_ROM_BYTES are read
 * (w_32, w_16, dig_rev tables, sbr_*_twiddle_*, post_fft_tbl, dct23_tw, qmf_c). */
#define XAAC_B200_QMF_ROM_BYTES 3464
int32_t xaac_b200_set_qmf_rom(xaac_b200_ctx *ctx, const void *tables, size_t bytes);

/* Complex ("HQ") 64-band QMF synthesis, batched.  Replaces ixheaacd_cplx_synt_qmffilt
 * (decoder/ixheaacd_qmf_dec.c:811-1129) as ixheaacd_sbr_dec calls it for a non-PS, non-low-power channel
 * (decoder/ixheaacd_sbr_dec.c:1273), together with the link-time leaves ixheaacd_inv_emodulation / cos_sin_mod /
 * radix4bfly / postradixcompute2 / shiftrountine_with_rnd / sbr_qmfsyn64_winadd
 * (decoder/generic/ixheaacd_qmf_dec_generic.c) and the selector leaf ixheaacd_adjust_scale.
 * Unit = one frame of one output channel.
 *   matrix         [n_units][32][128] WORD32: per time slot re[64] | im[64] (the reference's qmf_real[i]/qmf_imag[i]
 *                  rows, slot stride 128 words); read-only
 *   filter_states  [n_units][1280] WORD16 (ia_sbr_qmf_filter_bank_struct.filter_states), in/out
 *   pos            [n_units][2] WORD16 {ixheaacd_drc_offset, filter_pos_syn - p_filter}, in/out
 *   params         [n_units][8] WORD16 {ov_lb_scale, lb_scale, hb_scale, st_syn_scale (ia_sbr_scale_fact_struct),
 *                  lsb, usb (filter bank), split (= op_delay, 6), 0}
 *   pcm            PCM16, 2048 samples per unit; ch_fac as for the IMDCT stage
 */
int32_t xaac_b200_qmf_synth_hq_dev(xaac_b200_ctx *ctx, const int32_t *d_matrix, int16_t *d_filter_states,
                                   int16_t *d_pos, const int16_t *d_params, int16_t *d_pcm, int64_t n_units,
                                   int32_t ch_fac, void *stream);

/* Device-resident synthesis-bank state for the host-buffer entry point (reset state: zero filter states,
 * offsets 0 — decoder/ixheaacd_sbrdec_initfuncs.c:1154-1213). */
typedef struct xaac_b200_qmf_synth_state xaac_b200_qmf_synth_state;
int32_t xaac_b200_qmf_synth_state_create(xaac_b200_ctx *ctx, int64_t n_units, xaac_b200_qmf_synth_state **state);
void xaac_b200_qmf_synth_state_destroy(xaac_b200_ctx *ctx, xaac_b200_qmf_synth_state *state);
int32_t xaac_b200_qmf_synth_state_upload(xaac_b200_ctx *ctx, xaac_b200_qmf_synth_state *state,
                                         const int16_t *filter_states, const int16_t *pos);
int32_t xaac_b200_qmf_synth_state_download(xaac_b200_ctx *ctx, xaac_b200_qmf_synth_state *state,
                                           int16_t *filter_states, int16_t *pos);
int32_t xaac_b200_qmf_synth_hq_host(xaac_b200_ctx *ctx, xaac_b200_qmf_synth_state *state, const int32_t *matrix,
                                    const int16_t *params, int16_t *pcm, int32_t ch_fac);

/* Complex ("HQ") 32-band QMF analysis, batched.  Replaces ixheaacd_cplx_anal_qmffilt
 * (decoder/generic/ixheaacd_qmf_dec_generic.c:590-741; prototype decoder/ixheaacd_qmf_dec.h:74-80) as
 * ixheaacd_sbr_dec calls it with low_pow_flag = 0 (decoder/ixheaacd_sbr_dec.c:1025), together with the link-time
 * leaves ixheaacd_sbr_qmfanal32_winadd, ixheaacd_fwd_modulation, ixheaacd_cos_sin_mod, ixheaacd_radix4bfly,
 * ixheaacd_postradixcompute4.  The stage sets lb_scale = -8 (generic:635); callers use that constant.
 * Unit = one frame of one core channel.
 *   pcm     PCM16 core samples, 1024 per unit; ch_fac as for the IMDCT stage (time_sample_buf stride)
 *   states  [n_units][320] WORD16 anal_filter_states, in/out
 *   pos     [n_units][2] WORD16 {core_samples_buffer - anal_filter_states, filter_pos - analy_win_coeff}, in/out
 *   usb     [n_units] WORD16 qmf_bank->usb
 *   matrix  [n_units][32][128] WORD32: qmf_real[i][0..31] at row offset 0, qmf_imag[i][0..31] at row offset 64;
 *           the other 64 words of each row are left untouched (the reference does not write them either)
 */
int32_t xaac_b200_qmf_anal_hq_dev(xaac_b200_ctx *ctx, const int16_t *d_pcm, int16_t *d_states, int16_t *d_pos,
                                  const int16_t *d_usb, int32_t *d_matrix, int64_t n_units, int32_t ch_fac,
                                  void *stream);

/* ---- fixed-point complex ("HQ") SBR HF generator, batched --------------------------------------------
 * Replaces ixheaacd_hf_generator (decoder/ixheaacd_lpp_tran.c:956-1258; prototype decoder/ixheaacd_lpp_tran.h:72-80) as
 * ixheaacd_sbr_dec calls it (decoder/ixheaacd_sbr_dec.c:1169), with its leaves ixheaacd_invfilt_level_emphasis,
 * ixheaacd_filterstep3 and the selector leaves ixheaacd_covariance_matrix_calc_2 and ixheaacd_fix_div.
 * Unit = one frame of one SBR channel.
 *   lpc      [n_units][2][128] WORD32: hf_generator->lpc_filt_states_real[i] (64) | _imag[i] (64); read-only
 *   matrix   [n_units][38][128] WORD32: qmf_real[i] (64) | qmf_imag[i] (64) for the 6 overlap + 32 current slots;
 *            bands >= max_qmf_subband are generated in place
 *   params   [n_units][80] WORD16, XAAC_HF_* offsets below: ia_transposer_settings_struct
 *            (decoder/ixheaacd_lpp_tran.h:50-57) followed by the scalar arguments of the reference call
 *   bw_prev  [n_units][6] WORD32 hf_generator->bw_array_prev, in/out
 *   hb_scale [n_units] WORD16 sbr_scale_factor->hb_scale produced by the stage */
#define XAAC_HF_NUM_PATCHES 0
#define XAAC_HF_START_PATCH 1
#define XAAC_HF_STOP_PATCH 2
#define XAAC_HF_NUM_COLUMNS 3
#define XAAC_HF_BW_BORDERS 4       /* [10] */
#define XAAC_HF_PATCH 14           /* [6][6] src_start, src_end, guard_start, dst_start, dst_end, num_bands */
#define XAAC_HF_FACTOR 50          /* time_step */
#define XAAC_HF_NUM_IF_BANDS 51
#define XAAC_HF_START_IDX 52       /* border_vec[0] */
#define XAAC_HF_STOP_IDX 53        /* border_vec[num_env] - num_time_slots */
#define XAAC_HF_INVF 54            /* [10] sbr_invf_mode */
#define XAAC_HF_INVF_PREV 64       /* [10] sbr_invf_mode_prev */
#define XAAC_HF_OV_LB_SCALE 74
#define XAAC_HF_LB_SCALE 75
#define XAAC_HF_MAX_QMF_SUBBAND 76
#define XAAC_HF_PARAM_WORDS 80
int32_t xaac_b200_hf_generator_hq_dev(xaac_b200_ctx *ctx, const int32_t *d_lpc, int32_t *d_matrix,
                                      const int16_t *d_params, int32_t *d_bw_prev, int16_t *d_hb_scale,
                                      int64_t n_units, void *stream);

/* ---- fixed-point complex ("HQ") SBR envelope adjuster, batched -----------------------------------------
 * ROM: `env_tables` points at the host's ia_env_calc_tables_struct (decoder/ixheaacd_sbr_rom.h:59-68; what the
 * reference reaches as ptr_sbr_tables->env_calc_tables_ptr / ->sbr_rand_ph), `misc_tables` at the host's
 * ixheaacd_misc_tables (decoder/ixheaacd_common_rom.h:27-44; pstr_common_tables) of which the leading
 * XAAC_B200_MISC_ROM_BYTES (through inv_table and sqrt_table) are read. */
#define XAAC_B200_ENV_ROM_BYTES 2404
#define XAAC_B200_MISC_ROM_BYTES 2470
int32_t xaac_b200_set_env_rom(xaac_b200_ctx *ctx, const void *env_tables, size_t env_bytes, const void *misc_tables,
                              size_t misc_bytes);

/* ia_sbr_scale_fact_struct (decoder/ixheaacd_sbr_scale.h:23-31) as WORD16[8] */
#define XAAC_SF_LB 0
#define XAAC_SF_ST_LB 1
#define XAAC_SF_OV_LB 2
#define XAAC_SF_HB 3
#define XAAC_SF_OV_HB 4
#define XAAC_SF_ST_SYN 5
#define XAAC_SF_PS 6
/* Per-frame SBR side-info record, WORD16[XAAC_ENV_PARAM_WORDS]: the fields of ia_sbr_header_data_struct,
 * ia_freq_band_data_struct (decoder/ixheaacd_env_extr_part.h:33-100), ia_frame_info_struct and
 * ia_sbr_frame_info_data_struct / ia_sbr_prev_frame_data_struct (decoder/ixheaacd_env_extr.h:44-120) the stage reads. */
#define XAAC_ENV_NUM_TIME_SLOTS 0
#define XAAC_ENV_TIME_STEP 1
#define XAAC_ENV_CHANNEL_MODE 2        /* 1 SBR_MONO, 2 SBR_STEREO, 3 PS_STEREO */
#define XAAC_ENV_LIMITER_GAINS 3
#define XAAC_ENV_INTERPOL_FREQ 4
#define XAAC_ENV_SMOOTHING_MODE 5
#define XAAC_ENV_NUM_SF_LO 6
#define XAAC_ENV_NUM_SF_HI 7
#define XAAC_ENV_NUM_NF_BANDS 8
#define XAAC_ENV_SUB_BAND_START 9
#define XAAC_ENV_SUB_BAND_END 10
#define XAAC_ENV_NUM_LF_BANDS 11
#define XAAC_ENV_NUM_ENV 12
#define XAAC_ENV_TRANSIENT_ENV 13
#define XAAC_ENV_MAX_QMF_SUBBAND 14      /* frame_data->max_qmf_subband_aac */
#define XAAC_ENV_MAX_QMF_SUBBAND_PREV 15 /* frame_data_prev->max_qmf_subband_aac */
#define XAAC_ENV_BORDER_VEC 16           /* [9]  */
#define XAAC_ENV_FREQ_RES 25             /* [8]  */
#define XAAC_ENV_NOISE_BORDER_VEC 33     /* [3]  */
#define XAAC_ENV_LIM_TBL 36              /* [13] freq_band_tbl_lim */
#define XAAC_ENV_FREQ_LO 49              /* [29] freq_band_tbl_lo */
#define XAAC_ENV_FREQ_HI 78              /* [57] freq_band_tbl_hi */
#define XAAC_ENV_FREQ_NOISE 135          /* [6]  freq_band_tbl_noise */
#define XAAC_ENV_NOISE_FLOOR 141         /* [10] int_noise_floor */
#define XAAC_ENV_ADD_HARMONICS 151       /* [56] add_harmonics */
#define XAAC_ENV_SF_ARR 207              /* [448] int_env_sf_arr */
#define XAAC_ENV_PARAM_WORDS 656
/* ia_sbr_calc_env_struct (decoder/ixheaacd_env_calc.h:24-33) as WORD16[XAAC_ENV_STATE_WORDS] */
#define XAAC_ENV_ST_FILT_ME 0        /* [112] filt_buf_me (mantissa, exponent pairs) */
#define XAAC_ENV_ST_FILT_NOISE 112   /* [56]  filt_buf_noise_m */
#define XAAC_ENV_ST_NOISE_E 168
#define XAAC_ENV_ST_START_UP 169
#define XAAC_ENV_ST_PH_INDEX 170
#define XAAC_ENV_ST_TRANS_PREV 171
#define XAAC_ENV_ST_HARM_INDEX 172
#define XAAC_ENV_ST_HARM_PREV 173    /* [56]  harm_flags_prev */
#define XAAC_ENV_STATE_WORDS 232
/* Replaces ixheaacd_calc_sbrenvelope (decoder/ixheaacd_env_calc.c:692-1015; prototype decoder/ixheaacd_env_calc.h:35-45)
 * as ixheaacd_sbr_dec calls it with low_pow_flag = 0 (decoder/ixheaacd_sbr_dec.c:1195), with ixheaacd_adj_timeslot
 * (decoder/ixheaacd_env_dec.c:845) and the selector leaves ixheaacd_enery_calc_per_subband, ixheaacd_conv_ergtoamplitude,
 * ixheaacd_ixheaacd_expsubbandsamples, ixheaacd_adjust_scale.  Unit = one frame of one SBR channel.
 *   params  [n_units][656] WORD16 side info (read-only)
 *   sf      [n_units][8]   WORD16 scale factors; reads lb/hb/ov_hb, writes hb_scale and ov_hb_scale
 *   state   [n_units][232] WORD16 adjuster state, in/out
 *   matrix  [n_units][38][128] WORD32 QMF rows (re[64] | im[64]); bands >= max_qmf_subband adjusted in place
 *   err     [n_units] WORD32 or NULL: the reference's per-call return value (0 or 0x80000000) */
int32_t xaac_b200_calc_sbrenvelope_hq_dev(xaac_b200_ctx *ctx, const int16_t *d_params, int16_t *d_sf, int16_t *d_state,
                                          int32_t *d_matrix, int32_t *d_err, int64_t n_units, void *stream);

/* ---- whole fixed-point HQ SBR stage, batched ------------------------------------------------------------
 * Replaces ixheaacd_sbr_dec (decoder/ixheaacd_sbr_dec.c:662-1310, fixed-point branch, low_pow_flag = 0, 1024-sample
 * core frames, non-ELD/LD object types, no DRC/MPS) as ixheaacd_applysbr calls it (decoder/ixheaacd_sbrdecoder.c:877):
 * overlap hand-over + ixheaacd_rescale_x_overlap, ixheaacd_cplx_anal_qmffilt, the ixheaacd_expsubbandsamples /
 * ixheaacd_adjust_scale bookkeeping, ixheaacd_hf_generator, ixheaacd_calc_sbrenvelope, LPC / overlap state update,
 * [ixheaacd_init_ps_scale + per-slot ixheaacd_init_rot_env / ixheaacd_apply_ps] and ixheaacd_cplx_synt_qmffilt (x2 with PS).
 * Unit = one frame of one SBR channel (with PS: one mono core channel in, stereo out).
 *
 * ROM: in addition to the QMF / envelope ROMs, `ps_tables` points at the host's ia_ps_tables_struct
 * (decoder/ixheaacd_sbr_rom.h:177-238; sbr_tables_ptr->ps_tables_ptr), leading XAAC_B200_PS_ROM_BYTES read. */
#define XAAC_B200_PS_ROM_BYTES 1230
int32_t xaac_b200_set_ps_rom(xaac_b200_ctx *ctx, const void *ps_tables, size_t bytes);

/* Per-frame side-info record of the stage, WORD16[XAAC_SIDE_WORDS] */
#define XAAC_SIDE_ENV 0        /* [656] XAAC_ENV_* (MAX_QMF_SUBBAND_PREV is taken from the state) */
#define XAAC_SIDE_HF 656       /* [80]  XAAC_HF_*: transposer settings, NUM_IF_BANDS and INVF are read, the rest is derived */
#define XAAC_SIDE_APPLY 736    /* apply_processing argument (sync_state == SBR_ACTIVE) */
#define XAAC_SIDE_PS 737       /* 0 no PS this frame; 1 PS, stereo rotation exactly as the reference's own x86-64 gcc build
                                  executes ixheaacd_apply_rot_dec (its type-punned H11_H12 copy is read back as zero, so QMF
                                  bands 3..usb-1 of both outputs are 0 - bit-exact with that build); 2 PS, rotation as the C
                                  source is written (bit-exact with the same file built with -fno-strict-aliasing) */
#define XAAC_SIDE_PS_PRM 744   /* [488] XAAC_PS_PRM_*: ia_ps_dec_struct.iid_quant, num_env, border_position[7],
                                  iid_par_table[7][34], icc_par_table[7][34] after ixheaacd_decode_ps_data */
#define XAAC_SIDE_WORDS 1232
#define XAAC_ENV_PRM_WORDS 656
#define XAAC_ENV_ST_WORDS 232
#define XAAC_PS_PRM_IID_QUANT 0
#define XAAC_PS_PRM_NUM_ENV 1
#define XAAC_PS_PRM_BORDER 2
#define XAAC_PS_PRM_IID 9
#define XAAC_PS_PRM_ICC 247
/* Channel state as a host blob, WORD16[XAAC_SBR_ST_WORDS] per unit (32-bit members at even offsets): what
 * ia_sbr_dec_struct / ia_sbr_prev_frame_data_struct carry from frame to frame on this path */
#define XAAC_SBR_ST_ANAL_STATES 0   /* [320]  str_codec_qmf_bank.anal_filter_states */
#define XAAC_SBR_ST_ANAL_POS 320    /* [2]    {core_samples_buffer - anal_filter_states, filter_pos - qmf_c} */
#define XAAC_SBR_ST_SYN_POS 322     /* [2]    {ixheaacd_drc_offset, filter_pos_syn - qmf_c} */
#define XAAC_SBR_ST_SF 324          /* [8]    str_sbr_scale_fact (XAAC_SF_*) */
#define XAAC_SBR_ST_MISC 332        /* [16]   0 prev max_qmf_subband_aac, 1 prev end_position, 2..11 prev sbr_invf_mode,
                                              12 codec bank usb, 13 synthesis bank lsb, 14 synthesis bank usb */
#define XAAC_SBR_MISC_MAX_QMF_PREV 0 /* words of XAAC_SBR_ST_MISC */
#define XAAC_SBR_MISC_END_POS_PREV 1
#define XAAC_SBR_MISC_INVF_PREV 2    /* [10] */
#define XAAC_SBR_MISC_CODEC_USB 12
#define XAAC_SBR_MISC_SYN_LSB 13
#define XAAC_SBR_MISC_SYN_USB 14
#define XAAC_SBR_ST_ENV 348         /* [232]  str_sbr_calc_env (XAAC_ENV_ST_*) */
#define XAAC_SBR_ST_SYN_STATES 580  /* [1280] str_synthesis_qmf_bank.filter_states */
#define XAAC_SBR_ST_BW_PREV 1860    /* WORD32[6]      str_hf_generator.bw_array_prev */
#define XAAC_SBR_ST_LPC 1872        /* WORD32[2][128] lpc_filt_states_real[i] (64, 32 used) | _imag[i] */
#define XAAC_SBR_ST_OV 2384         /* WORD32[6][128] ptr_sbr_overlap_buf */
#define XAAC_SBR_ST_WORDS 3920
/* PS state as a host blob, WORD16[XAAC_PS_ST_WORDS] per unit: ia_ps_dec_struct (decoder/ixheaacd_ps_dec.h:97-141)
 * and the right channel's synthesis bank + scale factors */
#define XAAC_PS_ST_AP 0
#define XAAC_PS_ST_LD 128
#define XAAC_PS_ST_SD 464
#define XAAC_PS_ST_SER 528
#define XAAC_PS_ST_SUB 1488
#define XAAC_PS_ST_SUB_SER 1552
#define XAAC_PS_ST_HVEC 2032
#define XAAC_PS_ST_IDX 2320         /* [12] delay_buf_idx_ser[3], delay_buf_idx, delay_buf_idx_long, delay_buffer_scale,
                                            usb, -, right bank lsb, right bank usb */
#define XAAC_PS_IDX_SER 0            /* words of XAAC_PS_ST_IDX: [3] delay_buf_idx_ser */
#define XAAC_PS_IDX_DELAY 3
#define XAAC_PS_IDX_DELAY_LONG 4
#define XAAC_PS_IDX_SCALE 5
#define XAAC_PS_IDX_USB 6
#define XAAC_PS_IDX_LSB_R 8
#define XAAC_PS_IDX_USB_R 9
#define XAAC_PS_ST_PEAK 2332        /* WORD32[3][20] */
#define XAAC_PS_ST_HYB 2452         /* WORD32[3][2][12] */
#define XAAC_PS_ST_SYN_STATES_R 2596
#define XAAC_PS_ST_SYN_POS_R 3876
#define XAAC_PS_ST_SF_R 3878
#define XAAC_PS_ST_WORDS 3888

/* Device-resident state of n_units channels (structure of arrays in HBM) plus the stage's scratch QMF matrices.
 * upload / download move host blobs (checkpoint / resume, or hand-over from / to the CPU reference at a frame
 * boundary); either blob pointer may be NULL.  A new state is all-zero; upload the decoder's reset state
 * (decoder/ixheaacd_sbrdec_initfuncs.c:599-1213) before the first frame. */
typedef struct xaac_b200_sbr_state xaac_b200_sbr_state;
/* with_ps: 0 = one HQ channel per unit, 1 = HQ channel + parametric stereo (second synthesis bank),
 * XAAC_B200_SBR_STATE_LP = low-power channel (xaac_b200_sbr_dec_lp_dev only; no stage scratch is allocated) */
#define XAAC_B200_SBR_STATE_LP 2
int32_t xaac_b200_sbr_state_create(xaac_b200_ctx *ctx, int64_t n_units, int32_t with_ps, xaac_b200_sbr_state **state);
void xaac_b200_sbr_state_destroy(xaac_b200_ctx *ctx, xaac_b200_sbr_state *state);
int32_t xaac_b200_sbr_state_upload(xaac_b200_ctx *ctx, xaac_b200_sbr_state *state, const int16_t *st_blob,
                                   const int16_t *ps_blob);
int32_t xaac_b200_sbr_state_download(xaac_b200_ctx *ctx, xaac_b200_sbr_state *state, int16_t *st_blob,
                                     int16_t *ps_blob);
/* One frame for every unit of `state` (asynchronous on `stream`):
 *   d_side     [n][1232] WORD16 side info
 *   d_time_in  [n][1024] PCM16 core-coder output (after ixheaacd_allocate_sbr_scr's WORD32 -> WORD16 conversion)
 *   d_time_out state without PS: [n][2048] PCM16; state with PS: [n][2048][2] interleaved L/R like the reference's
 *              stereo time buffer (R is written only for units whose frame runs PS)
 *   d_err      [n] WORD32 or NULL: the stage's return value per unit (0 / 0x80000000); a unit that fails keeps neither a
 *              valid output nor a defined state, as in the reference, which aborts the frame */
int32_t xaac_b200_sbr_dec_hq_dev(xaac_b200_ctx *ctx, xaac_b200_sbr_state *state, const int16_t *d_side,
                                 const int16_t *d_time_in, int16_t *d_time_out, int32_t *d_err, void *stream);
/* The same stage fed with the core coder's WORD32 output: d_w32 [n][1024] and d_qshift_adj [n] as written by
 * xaac_b200_imdct_process_dev.  The WORD32 -> WORD16 hand-over of ixheaacd_allocate_sbr_scr (decoder/ixheaacd_api.c:337-370,
 * round16(shl32_sat(x, qshift_adj))) happens in the analysis bank's load, so the PCM16 copy of the core output never
 * exists in HBM (one launch and 6 KB of traffic per unit less than xaac_b200_imdct_out_to_pcm16_dev + the call above). */
int32_t xaac_b200_sbr_dec_hq_w32_dev(xaac_b200_ctx *ctx, xaac_b200_sbr_state *state, const int16_t *d_side,
                                     const int32_t *d_w32, const int8_t *d_qshift_adj, int16_t *d_time_out,
                                     int32_t *d_err, void *stream);

/* ---- low-power (real-valued) SBR stage: ixheaacd_sbr_dec with low_pow_flag = 1 --------------------------------
 * The path the reference runs for stereo HE-AACv1 in its fixed-point mode (decoder/ixheaacd_sbrdecoder.c:408-419:
 * low_pow_flag = 1 unless the stream is mono / PS): 32-band real analysis (ixheaacd_cplx_anal_qmffilt with
 * ixheaacd_dct3_32, decoder/generic/ixheaacd_qmf_dec_generic.c:63-239, 590-741), ixheaacd_low_pow_hf_generator
 * (decoder/ixheaacd_lpp_tran.c:843-954), ixheaacd_calc_sbrenvelope with the low-power leaves and alias reduction
 * (decoder/ixheaacd_env_calc.c:78-227, 692-1015, 1564-1757) and the real 64-band synthesis (ixheaacd_cplx_synt_qmffilt
 * with ixheaacd_inv_modulation_lp / ixheaacd_dct2_64, decoder/ixheaacd_qmf_dec.c:72-211, 811-1129).
 * ONE fused kernel: the unit's QMF matrix never leaves shared memory.  Same side-info record and state blob as the HQ
 * stage (the overlap slots hold 6 real rows of 64 words, the LPC rows 32 real words each); the state must have been
 * created with XAAC_B200_SBR_STATE_LP or with_ps = 0.
 *   d_side     [n][1232] WORD16 side info (PS part ignored)
 *   d_time_in  [n][1024] PCM16 core-coder output of the unit's channel
 *   d_time_out [n / out_ch][2048][out_ch] PCM16: unit u is channel u % out_ch of frame u / out_ch (out_ch = 2 gives the
 *              reference's interleaved stereo time buffer)
 *   d_err      [n] WORD32 or NULL (0 / 0x80000000) */
int32_t xaac_b200_sbr_dec_lp_dev(xaac_b200_ctx *ctx, xaac_b200_sbr_state *state, const int16_t *d_side,
                                 const int16_t *d_time_in, int16_t *d_time_out, int32_t out_ch, int32_t *d_err,
                                 void *stream);

/* The low-power stage fed with the core coder's WORD32 output (d_w32 [n][1024], d_qshift_adj [n] as written by
 * xaac_b200_imdct_process_dev): the hand-over of ixheaacd_allocate_sbr_scr (decoder/ixheaacd_api.c:337-370) runs in the kernel's load. */
int32_t xaac_b200_sbr_dec_lp_w32_dev(xaac_b200_ctx *ctx, xaac_b200_sbr_state *state, const int16_t *d_side,
                                     const int32_t *d_w32, const int8_t *d_qshift_adj, int16_t *d_time_out, int32_t out_ch,
                                     int32_t *d_err, void *stream);

/* Host-buffer entry point for a whole HE-AAC (v1 mono / v2) frame per unit: IMDCT + window/OLA of the core channel
 * (xaac_b200_imdct_process_dev), the WORD32 -> PCM16 hand-over (inside the analysis bank's load, as in
 * xaac_b200_sbr_dec_hq_w32_dev) and the SBR stage, chunked and pipelined over internal streams (H2D, kernels, D2H overlap).  Both states stay resident in HBM.
 *   spec [n][1024] WORD32, ics [n][2], side [n][1232] WORD16 (host, pinned recommended)
 *   pcm  [n][2048] PCM16, or [n][2048][2] for a PS state; err [n] WORD32 or NULL */
int32_t xaac_b200_heaac_frame_host(xaac_b200_ctx *ctx, xaac_b200_imdct_state *imdct_state, xaac_b200_sbr_state *sbr_state,
                                   const int32_t *spec, const uint8_t *ics, const int16_t *side, int16_t *pcm,
                                   int32_t *err);

/* The same for stereo HE-AACv1 streams on the low-power path: unit = one core channel (units 2k / 2k+1 = L / R of stream
 * k with out_ch = 2), `sbr_state` created with XAAC_B200_SBR_STATE_LP.  pcm [n / out_ch][2048][out_ch] PCM16. */
int32_t xaac_b200_heaac_lp_frame_host(xaac_b200_ctx *ctx, xaac_b200_imdct_state *imdct_state,
                                      xaac_b200_sbr_state *sbr_state, const int32_t *spec, const uint8_t *ics,
                                      const int16_t *side, int16_t *pcm, int32_t out_ch, int32_t *err);

/* ---- stage glue: WORD32 IMDCT output -> PCM16 ---------------------------------------------
 * mode 0 replaces the conversion loop of ixheaacd_allocate_sbr_scr (decoder/ixheaacd_api.c:337-370):
 *        round16(shl32_sat(x, qshift_adj)), the core-coder -> SBR hand-over;
 * mode 1 replaces ixheaacd_scale_adjust (decoder/ixheaacd_peak_limiter.c:324-333) + the round16 loop of
 *        ixheaacd_dec_execute (decoder/ixheaacd_api.c:3676-3681): the AAC-LC output stage with -peak_limiter_off:1.
 *   d_in [n_units][1024] WORD32, d_qshift_adj [n_units] (as written by xaac_b200_imdct_process_dev), d_out [n_units][1024] */
int32_t xaac_b200_imdct_out_to_pcm16_dev(xaac_b200_ctx *ctx, const int32_t *d_in, const int8_t *d_qshift_adj,
                                         int16_t *d_out, int64_t n_units, int32_t mode, void *stream);

/* ---- AAC-LC output stage: peak limiter + round16 (SURVEY.md 8a-F "LC output") -----------------------------------------
 * Batched drop-in for ixheaacd_peak_limiter_process(ia_peak_limiter_struct *, VOID *samples, UWORD32 frame_len,
 * UWORD8 *qshift_adj) (decoder/ixheaacd_peak_limiter.c:177-307) followed by the round16 loop of ixheaacd_dec_execute
 * (decoder/ixheaacd_api.c:3676-3681): what the reference runs on the WORD32 IMDCT output of an AAC-LC stream with its
 * default flags (-peak_limiter_off:0).  Unit = one stream (1 or 2 channels, 1024 samples per channel).
 * Per-stream state: ia_peak_limiter_struct (decoder/ixheaacd_peak_limiter_struct_def.h:29-48) as 32-bit words: */
#define XAAC_PL_ATTACK_CONST 0   /* float  attack_constant */
#define XAAC_PL_RELEASE_CONST 1  /* float  release_constant */
#define XAAC_PL_GAIN_MOD 2       /* float  gain_modified */
#define XAAC_PL_MIN_GAIN 3       /* float  min_gain (written by the stage) */
#define XAAC_PL_PSG 4            /* double pre_smoothed_gain */
#define XAAC_PL_ATTACK 6         /* attack_time_samples (<= 512) */
#define XAAC_PL_DELAY_IDX 7      /* delayed_input_index */
#define XAAC_PL_MAX_IDX 8        /* max_idx */
#define XAAC_PL_CIR 9            /* cir_buf_pnt */
#define XAAC_PL_LIMITER_ON 10
#define XAAC_PL_NUM_CH 11
#define XAAC_PL_MAX_BUF 12       /* float[512]    max_buf */
#define XAAC_PL_DELAYED 524      /* float[512][2] delayed_input */
#define XAAC_PL_STATE_WORDS 1548
/* host-side: the state ixheaacd_peak_limiter_init produces (decoder/ixheaacd_peak_limiter.c:45-75) */
int32_t xaac_b200_peak_limiter_state_init(int32_t *state, int32_t num_channels, int32_t sample_rate);
/*   d_state      [n][1548] in/out
 *   d_samples    [n][1024][ch] WORD32, interleaved like the reference's time_data (e.g. xaac_b200_imdct_process_dev, ch_fac = ch)
 *   d_qshift_adj [n][ch] (as written by xaac_b200_imdct_process_dev)
 *   d_out32      [n][1024][ch] limited WORD32 samples, or NULL;  d_pcm16 [n][1024][ch] PCM16, or NULL (one of them required)
 *   d_err        [n] or NULL: 0 / 0x80000000 (state not supported: more than 2 channels, attack > 512, limiter off) */
int32_t xaac_b200_peak_limiter_dev(xaac_b200_ctx *ctx, int32_t *d_state, const int32_t *d_samples,
                                   const int8_t *d_qshift_adj, int32_t *d_out32, int16_t *d_pcm16, int32_t *d_err,
                                   int64_t n_units, int32_t num_channels, void *stream);

/* ---- USAC frequency-domain core transform -------------------------------------------------------
 * Batched drop-in for ixheaacd_fd_frm_dec(ia_usac_data_struct *, WORD32 i_ch) (decoder/ixheaacd_imdct.c:596; called from
 * ixheaacd_core_coder_data, decoder/ixheaacd_ext_ch_ele.c:991) for pure frequency-domain streams: previous frame FD
 * (td_frame_prev = 0), no FAC data, frame ok, ccfl = 1024.  Covers ixheaacd_fd_imdct_long / _short, ixheaacd_acelp_imdct,
 * the saturating FFT ixheaacd_complex_fft_p2_dec (decoder/ixheaacd_fft.c:1412; selector entry ixheaacd_complex_fft_p2,
 * also ixheaacd_calc_pre_twid / ixheaacd_calc_post_twid) and the windowing leaves of decoder/ixheaacd_basic_ops.c.
 * LPD / FAC transitions stay on the host (they need the ACELP state).
 * ROM blob (XAAC_UROM_BYTES): the reference's tables, concatenated in this order — in a drop-in deployment the host
 * memcpy's them from its own const arrays:
 *   ixheaacd_twiddle_table_fft_32x32[514], ixheaacd_pre_post_twid_cos_512[512], _sin_512[512], _cos_64[64], _sin_64[64],
 *   ixheaacd_sine_win_1024[1024], ixheaacd_kbd_win1024[1024], ixheaacd_sine_win_128[128], ixheaacd_kbd_win128[128] (WORD32). */
#define XAAC_UROM_FFT_TW 0
#define XAAC_UROM_COS512 2056
#define XAAC_UROM_SIN512 4104
#define XAAC_UROM_COS64 6152
#define XAAC_UROM_SIN64 6408
#define XAAC_UROM_SINE1024 6664
#define XAAC_UROM_KBD1024 10760
#define XAAC_UROM_SINE128 14856
#define XAAC_UROM_KBD128 15368
#define XAAC_UROM_BYTES 15880
int32_t xaac_b200_set_usac_rom(xaac_b200_ctx *ctx, const void *tables, size_t bytes);
/*   d_coef    [n][1024] WORD32 usac_data->coef_fix[ch] (read-only here; the reference uses it as workspace)
 *   d_overlap [n][1024] WORD32 usac_data->overlap_data_ptr[ch], in/out
 *   d_wstate  [n] uint8 usac_data->window_shape_prev[ch], in/out (set to this frame's shape, ext_ch_ele.c:968)
 *   d_ics     [n][2] uint8 {window_sequence (0 ONLY_LONG, 1 LONG_START, 2 EIGHT_SHORT, 3 LONG_STOP, 4 STOP_START), window_shape}
 *   d_out     [n][1024] WORD32 usac_data->output_data_ptr[ch] (Q15; the reference then scales by 2^-15 to float) */
int32_t xaac_b200_usac_fd_frm_dec_dev(xaac_b200_ctx *ctx, const int32_t *d_coef, int32_t *d_overlap, uint8_t *d_wstate,
                                      const uint8_t *d_ics, int32_t *d_out, int64_t n_units, void *stream);

/* ---- eSBR 64-band QMF synthesis bank (first piece of SURVEY.md 8a-E, the float eSBR path) ----------------------------
 * Batched per-slot core of ixheaacd_esbr_synthesis_filt_block (decoder/ixheaacd_sbr_dec.c:447, lines 583-654:
 * stereo_config_idx <= 0, 64 synthesis channels, 32 time slots): float -> WORD32 (x 64), ixheaacd_esbr_inv_modulation
 * (decoder/ixheaacd_qmf_dec.c:733; link-time leaves ixheaacd_esbr_cos_sin_mod, ixheaacd_esbr_radix4bfly,
 * ixheaacd_esbr_postradixcompute2, decoder/generic/ixheaacd_qmf_dec_generic.c), ixheaacd_shiftrountine_with_rnd_hq,
 * ixheaacd_esbr_qmfsyn64_winadd, WORD32 -> float.  Integer arithmetic inside, hence bit-exact float output.  The
 * regrouping, PS and DRC branches of the stage function stay with the caller.
 * ROM blob (XAAC_EROM_BYTES), members of ia_qmf_dec_tables_struct (decoder/ixheaacd_sbr_rom.h:96-105) concatenated:
 *   esbr_qmf_c[1280], esbr_w_32[60], esbr_sin_cos_twiddle_l64[64], esbr_alt_sin_twiddle_l64[32], esbr_w_16[24],
 *   esbr_sin_cos_twiddle_l32[32], esbr_alt_sin_twiddle_l32[16], esbr_t_cos_sin_l32[64] (WORD32). */
#define XAAC_EROM_QMF_C 0
#define XAAC_EROM_W32 5120
#define XAAC_EROM_SINCOS_L64 5360
#define XAAC_EROM_ALTSIN_L64 5616
#define XAAC_EROM_W16 5744         /* esbr_w_16[24] */
#define XAAC_EROM_SINCOS_L32 5840  /* esbr_sin_cos_twiddle_l32[32] */
#define XAAC_EROM_ALTSIN_L32 5968  /* esbr_alt_sin_twiddle_l32[16] */
#define XAAC_EROM_TCOS_L32 6032    /* esbr_t_cos_sin_l32[64] */
#define XAAC_EROM_BYTES 6288
int32_t xaac_b200_set_esbr_rom(xaac_b200_ctx *ctx, const void *tables, size_t bytes);
/*   d_qmf    [n][32][128] float: per slot qmf_buf_real[i][0..63] | qmf_buf_imag[i][0..63]
 *   d_states [n][1280] WORD32 ia_sbr_qmf_filter_bank_struct.filter_states_32, in/out
 *   d_pos    [n][2] WORD32 {ixheaacd_drc_offset, filter_pos_syn_32 - esbr_qmf_c}, in/out
 *   d_out    [n][2048] float time samples (ptr_sbr_dec->time_sample_buf)
 *   d_err    [n] or NULL: 0 / 0x80000000 (ring positions the reference cannot produce) */
int32_t xaac_b200_esbr_synth64_dev(xaac_b200_ctx *ctx, const float *d_qmf, int32_t *d_states, int32_t *d_pos, float *d_out,
                                   int32_t *d_err, int64_t n_units, void *stream);

/* The same bank with ixheaacd_samples_sat (decoder/ixheaacd_decode_main.c:82-104, pcmsize 16) fused into its store: clamp to
 * [-32768, 32767], C cast (truncation towards zero), channel-interleaved.  Unit u is channel u % ch_fac of stream u / ch_fac;
 * its sample i goes to d_pcm16[(stream * 2048 + i) * ch_fac + channel].  d_out may be NULL (PCM only). */
int32_t xaac_b200_esbr_synth64_pcm16_dev(xaac_b200_ctx *ctx, const float *d_qmf, int32_t *d_states, int32_t *d_pos,
                                         float *d_out, int16_t *d_pcm16, int32_t ch_fac, int32_t *d_err, int64_t n_units,
                                         void *stream);

/* eSBR 32-band QMF analysis bank: batched ixheaacd_esbr_analysis_filt_block(ia_sbr_dec_struct *, ia_sbr_tables_struct *,
 * WORD32 op_delay) (decoder/ixheaacd_sbr_dec.c:185-295) for 32 analysis channels and 32 time slots, with its leaves
 * ixheaacd_esbr_qmfanal32_winadd (decoder/ixheaacd_qmf_dec.c:537), ixheaacd_esbr_fwd_modulation, ixheaacd_esbr_cos_sin_mod,
 * ixheaacd_esbr_radix4bfly, ixheaacd_esbr_postradixcompute4 (decoder/generic/ixheaacd_qmf_dec_generic.c).  Bit-exact floats.
 *   d_time_in [n][1024] float  ptr_sbr_dec->time_sample_buf (core coder output x 2^-15, ixheaacd_ext_ch_ele.c:1040-1046)
 *   d_states  [n][320] WORD32  str_codec_qmf_bank.anal_filter_states_32, in/out
 *   d_pos     [n][2] WORD32    {state_new_samples_pos_low_32 - anal_filter_states_32, filter_pos_32 - esbr_qmf_c}, in/out
 *   d_qmf     [n][32][128] float: qmf_buf_real[op_delay + i][0..31] at +0, qmf_buf_imag[..][0..31] at +64 of slot row i */
int32_t xaac_b200_esbr_anal32_dev(xaac_b200_ctx *ctx, const float *d_time_in, int32_t *d_states, int32_t *d_pos, float *d_qmf,
                                  int32_t *d_err, int64_t n_units, void *stream);

/* The same bank with the core -> eSBR hand-overs fused into its load:
 *   _core_dev:  d_core [n][1024] WORD32 = usac_data->output_data_ptr[ch] (the output of xaac_b200_usac_fd_frm_dec_dev);
 *               time_sample_vector = (FLOAT32)x * 2^-15 (decoder/ixheaacd_ext_ch_ele.c:1040-1046)
 *   _pcm16_dev: d_pcm16 = the legacy core decoder's interleaved WORD16 time_data; unit u is channel u % ch_fac of stream
 *               u / ch_fac and reads (FLOAT32)time_data[ch_fac * i + channel] (decoder/ixheaacd_api.c:3384-3437) */
int32_t xaac_b200_esbr_anal32_core_dev(xaac_b200_ctx *ctx, const int32_t *d_core, int32_t *d_states, int32_t *d_pos,
                                       float *d_qmf, int32_t *d_err, int64_t n_units, void *stream);
int32_t xaac_b200_esbr_anal32_pcm16_dev(xaac_b200_ctx *ctx, const int16_t *d_pcm16, int32_t ch_fac, int32_t *d_states,
                                        int32_t *d_pos, float *d_qmf, int32_t *d_err, int64_t n_units, void *stream);

/* eSBR float HF generator: batched ixheaacd_generate_hf (decoder/ixheaacd_sbrdec_lpfuncs.c:981-1359) with
 * ixheaacd_esbr_calc_co_variance (:781), ixheaacd_esbr_chirp_fac_calc (:832) and ixheaacd_pre_processing (:928-979, with
 * ixheaacd_polyfit / ixheaacd_gausssolve), for the 2:1 system (is_usf_4 == 0) without LD-MPS and error concealment (units
 * asking for those get err = -2 and are left to the reference).
 * Float results are bit-identical to the reference build's (one IEEE rounding per operation, same order); the one exception is
 * a pre-processing gain, whose log10 / pow run in double on the device like the reference's libm calls and can round to the
 * neighbouring float (1 ulp, ~1e-8 of the values; inside the +-1 LSB the float path is graded at).
 * QMF buffers are [n][XAAC_EHF_ROWS][64] floats and are the reference's own arrays from their FIRST row, i.e. row r is row
 * r - SBR_HF_ADJ_OFFSET of the pointers ixheaacd_sbr_dec passes (decoder/ixheaacd_sbr_dec.c:921-929):
 *   d_src_re/im  ptr_sbr_dec->qmf_buf_real / qmf_buf_imag                  (read: bands < f_master_tbl[0])
 *   d_pv_re/im   ptr_sbr_dec->ph_vocod_qmf_real / ph_vocod_qmf_imag, or both NULL (no harmonic transposer)
 *   d_dst_re/im  ptr_sbr_dec->sbr_qmf_out_real / sbr_qmf_out_imag, in/out: exactly the cells the reference writes
 *   d_par        [n][XAAC_EHF_PAR_WORDS] WORD32, word offsets below
 *   d_bw_prev    [n][6] float  ptr_frame_data->bw_array_prev, in/out
 *   d_patch_out  [n][8] WORD32 {patch_param.num_patches, patch_param.start_subband[0..6]} or NULL
 *   d_err        [n] or NULL: 0, -1 (the reference's own failure returns) or -2 (outside the supported subset) */
#define XAAC_EHF_ROWS 40
#define XAAC_EHF_NUM_MF 0         /* pstr_freq_band_data->num_mf_bands */
#define XAAC_EHF_NUM_IF 1         /* pstr_freq_band_data->num_nf_bands */
#define XAAC_EHF_SB_START 2       /* pstr_freq_band_data->sub_band_start */
#define XAAC_EHF_BORDER_FIRST 3   /* str_frame_info_details.border_vec[0] */
#define XAAC_EHF_BORDER_LAST 4    /* str_frame_info_details.border_vec[num_env] */
#define XAAC_EHF_HBE_FLAG 5       /* ptr_header_data->hbe_flag */
#define XAAC_EHF_PATCHING_MODE 6  /* ptr_frame_data->sbr_patching_mode */
#define XAAC_EHF_FS 7             /* ptr_header_data->out_sampling_freq */
#define XAAC_EHF_PRE_PROC 8       /* ptr_header_data->pre_proc_flag: 1 runs ixheaacd_pre_processing (lpfuncs.c:928-979) */
#define XAAC_EHF_USF4 9           /* ptr_header_data->is_usf_4 (must be 0) */
#define XAAC_EHF_MPS_SBR 10       /* ptr_frame_data->mps_sbr_flag */
#define XAAC_EHF_COV_COUNT 11     /* ptr_frame_data->cov_count */
#define XAAC_EHF_INVF 16          /* ptr_frame_data->sbr_invf_mode[0..4] */
#define XAAC_EHF_INVF_PREV 21     /* ptr_frame_data->sbr_invf_mode_prev[0..4] */
#define XAAC_EHF_INVF_TBL 26      /* pstr_freq_band_data->freq_band_tbl_noise[1..5] */
#define XAAC_EHF_FMASTER 32       /* pstr_freq_band_data->f_master_tbl[0..56] */
#define XAAC_EHF_PAR_WORDS 96
int32_t xaac_b200_esbr_generate_hf_dev(xaac_b200_ctx *ctx, const float *d_src_re, const float *d_src_im, const float *d_pv_re,
                                       const float *d_pv_im, float *d_dst_re, float *d_dst_im, const int32_t *d_par,
                                       float *d_bw_prev, int32_t *d_patch_out, int32_t *d_err, int64_t n_units, void *stream);

/* eSBR float envelope adjuster: batched ixheaacd_sbr_env_calc (decoder/ixheaacd_esbr_envcal.c:71-908), the ORIG_SBR branch
 * (:611-860) and epilogue, for the 2:1 system.  Not covered (err = -2, unit left to the reference): reset_flag / a change
 * of sbr_patching_mode (ixheaacd_createlimiterbands is control plane — the host passes lim_table / gate_mode), PVC,
 * LD-MPS, inter-TES (inter_temp_shape_mode != 0), error concealment.  Results are bit-identical to the reference build.
 *   d_re / d_im  [n][XAAC_EHF_ROWS][64] float  sbr_qmf_out_real / sbr_qmf_out_imag from their first row (the destination
 *                of xaac_b200_esbr_generate_hf_dev), adjusted in place
 *   d_ipar       [n][XAAC_EEC_IPAR_WORDS] WORD32, word offsets below; the words marked in/out are updated
 *   d_fpar       [n][XAAC_EEC_FPAR_WORDS] float: flt_env_sf_arr[448] | flt_noise_floor[10]
 *   d_state      [n][640] float: frame_data->e_gain[5][64] | noise_buf[5][64], in/out
 *   d_err        [n] or NULL: 0, -1 / 0x80000000 (the reference's own failure returns), -2 (unsupported subset) */
#define XAAC_EEC_SB_START 0         /* pstr_freq_band_data->sub_band_start */
#define XAAC_EEC_SB_END 1           /* pstr_freq_band_data->sub_band_end */
#define XAAC_EEC_NUM_ENV 2          /* str_frame_info_details.num_env */
#define XAAC_EEC_TRANS_ENV 3        /* str_frame_info_details.transient_env */
#define XAAC_EEC_SHORT_PREV 4       /* env_short_flag_prev (in/out) */
#define XAAC_EEC_NUM_NOISE_ENV 5    /* str_frame_info_details.num_noise_env */
#define XAAC_EEC_NUM_SF_LO 6        /* num_sf_bands[LOW] */
#define XAAC_EEC_NUM_SF_HI 7        /* num_sf_bands[HIGH] */
#define XAAC_EEC_NUM_NF 8           /* num_nf_bands */
#define XAAC_EEC_SMOOTHING_MODE 9   /* pstr_sbr_header->smoothing_mode */
#define XAAC_EEC_INTERPOL_FREQ 10   /* pstr_sbr_header->interpol_freq */
#define XAAC_EEC_LIMITER_BANDS 11   /* pstr_sbr_header->limiter_bands */
#define XAAC_EEC_LIMITER_GAINS 12   /* pstr_sbr_header->limiter_gains */
#define XAAC_EEC_HARM_INDEX 13      /* harm_index (in/out) */
#define XAAC_EEC_PHASE_INDEX 14     /* phase_index (in/out) */
#define XAAC_EEC_START_UP 15        /* pstr_sbr_header->esbr_start_up (in/out) */
#define XAAC_EEC_RESET 16           /* reset_flag: esbr_start_up = 1, phase_index = 0 (esbr_envcal.c:169-172); needs LIM_REBUILT */
#define XAAC_EEC_SBR_MODE 17        /* sbr_mode (must be ORIG_SBR = 1) */
#define XAAC_EEC_USF4 18            /* is_usf_4 (must be 0) */
#define XAAC_EEC_PATCHING_CHANGED 19 /* sbr_patching_mode != prev_sbr_patching_mode; needs LIM_REBUILT */
#define XAAC_EEC_LIM_REBUILT 20     /* 1 = the host has run ixheaacd_createlimiterbands (esbr_envcal.c:173-188, control plane) for this
                                     * reset / patching-change frame and LIM_TABLE / GATE_MODE carry the result; without it such a
                                     * frame returns -2.  With LPP patching the patch table comes from the HF generator of the same
                                     * frame: call xaac_b200_esbr_dec_front_dev, read `patch`, rebuild, call _back_dev */
#define XAAC_EEC_BORDER 24          /* border_vec[9] */
#define XAAC_EEC_FREQ_RES 33        /* freq_res[8] */
#define XAAC_EEC_NOISE_BORDER 41    /* noise_border_vec[3] */
#define XAAC_EEC_INTER_TES 44       /* inter_temp_shape_mode[8], 0..3; an envelope with a non-zero mode needs the low band */
#define XAAC_EEC_GATE_MODE 52       /* gate_mode[4] */
#define XAAC_EEC_LIM_TABLE 56       /* lim_table[4][13] */
#define XAAC_EEC_TBL_NOISE 108      /* freq_band_tbl_noise[6] */
#define XAAC_EEC_TBL_LO 116         /* freq_band_tbl_lo[29] */
#define XAAC_EEC_TBL_HI 148         /* freq_band_tbl_hi[57] */
#define XAAC_EEC_ADD_HARM 208       /* add_harmonics[56] */
#define XAAC_EEC_HARM_PREV 264      /* harm_flag_prev[64], one byte each (in/out) */
#define XAAC_EEC_IPAR_WORDS 288
#define XAAC_EEC_SFB_NRG 0
#define XAAC_EEC_NOISE_FLOOR 448
#define XAAC_EEC_FPAR_WORDS 464
#define XAAC_EEC_STATE_WORDS 640
/* random_phase = ixheaac_random_phase[512][2] (common/ixheaac_esbr_rom.c:437), 4096 bytes */
int32_t xaac_b200_set_esbr_envcalc_rom(xaac_b200_ctx *ctx, const void *random_phase, size_t bytes);
int32_t xaac_b200_esbr_env_calc_dev(xaac_b200_ctx *ctx, float *d_re, float *d_im, int32_t *d_ipar, const float *d_fpar,
                                    float *d_state, int32_t *d_err, int64_t n_units, void *stream);
/* The same with the low band at hand, for envelopes that use inter-TES (XAAC_EEC_INTER_TES + env != 0: ixheaacd_apply_inter_tes,
 * decoder/ixheaacd_esbr_envcal.c:1021-1096, called at :824-836): d_low_re / d_low_im [n][low_rows][64] = qmf_buf_real / imag from
 * their first row (low_rows 40, or 72 with the harmonic transposer's delayed core QMF).  Without them (NULL, or the entry above)
 * a frame with an inter-TES envelope returns -2.  The whole-stage entries always pass the low band. */
int32_t xaac_b200_esbr_env_calc_tes_dev(xaac_b200_ctx *ctx, float *d_re, float *d_im, const float *d_low_re, const float *d_low_im,
                                        int32_t low_rows, int32_t *d_ipar, const float *d_fpar, float *d_state, int32_t *d_err,
                                        int64_t n_units, void *stream);

/* QMF harmonic transposer: batched ixheaacd_qmf_hbe_apply (decoder/ixheaacd_hbe_trans.c:224-296) with
 * ixheaacd_real_synth_filt / ixheaacd_complex_anal_filt (decoder/ixheaacd_esbr_polyphase.c:157, 48),
 * ixheaacd_hbe_post_anal_process + prod2/3/4 + xprod2/3/4 (hbe_trans.c:298-1606) and the FFTs of common/ixheaac_esbr_fft.c,
 * for the 2:1 system (32 QMF columns per call): bank sizes 4 / 8 / 12 / 16 (FFT banks) and 20 (direct form; for this size the
 * reference's FFT pointers stay NULL and it re-initialises the instance inside every call, hbe_trans.c:240-248 — the bank
 * histories then start from zero each frame, reproduced).  Float results are bit-identical to the reference build's.
 * ROM blob = the reference's global float tables (common/ixheaac_esbr_rom.c) concatenated, float-word offsets: */
#define XAAC_HROM_WIN 0          /* ixheaac_sub_samp_qmf_window_coeff[1560] */
#define XAAC_HROM_SYNCOS 1560    /* ixheaac_synth_cos_table_kl_4[16] | _8[32] | _12[48] | _16[64] */
#define XAAC_HROM_ANACS 1720     /* ixheaac_analy_cos_sin_table_kl_8[32] | _16[64] | _24[96] | _32[128] */
#define XAAC_HROM_COSTRANS 2040  /* ixheaac_cos_table_trans_qmf[7][64] */
#define XAAC_HROM_FFTTW 2488     /* ixheaac_twiddle_table_fft_float[514] (+ 2 words of padding) */
#define XAAC_HROM_TW24 3004      /* ixheaac_twidle_tbl_24[32] */
#define XAAC_HROM_TW48 3036      /* ixheaac_twidle_tbl_48[64] */
#define XAAC_HROM_PVCOS 3100     /* ixheaac_phase_vocoder_cos_table[64] */
#define XAAC_HROM_PVSIN 3164     /* ixheaac_phase_vocoder_sin_table[64] */
#define XAAC_HROM_INTERP 3228    /* ixheaac_hbe_post_anal_proc_interp_coeff[4][2] */
#define XAAC_HROM_SELCASE 3236   /* ixheaac_sel_case[5][8] */
#define XAAC_HROM_XP2 3276       /* ixheaac_hbe_x_prod_cos_table_trans_2[512] */
#define XAAC_HROM_XP3 3788       /* ixheaac_hbe_x_prod_cos_table_trans_3[512] */
#define XAAC_HROM_XP4 4300       /* ixheaac_hbe_x_prod_cos_table_trans_4[512] */
#define XAAC_HROM_XP41 4812      /* ixheaac_hbe_x_prod_cos_table_trans_4_1[512] */
#define XAAC_HROM_SYN20 5324     /* ixheaac_synth_cos_table_kl_20[800] */
#define XAAC_HROM_ANA40 6124     /* ixheaac_analy_cos_sin_table_kl_40[3200] */
#define XAAC_HROM_WORDS 9324
int32_t xaac_b200_set_hbe_rom(xaac_b200_ctx *ctx, const void *tables, size_t bytes);
/* d_cfg [n][XAAC_HBE_CFG_WORDS] WORD32: the members of ia_esbr_hbe_txposer_struct that ixheaacd_qmf_hbe_data_reinit
 * (hbe_trans.c:103-222, control plane, stays on the host) derives from the frequency tables, plus the call's pitch_in_bins */
#define XAAC_HBE_SYNTH_SIZE 0
#define XAAC_HBE_K_START 1
#define XAAC_HBE_START_BAND 2
#define XAAC_HBE_END_BAND 3
#define XAAC_HBE_MAX_STRETCH 4
#define XAAC_HBE_PITCH 5         /* ptr_frame_data->pitch_in_bins: < 12 plain products, >= 12 cross products */
#define XAAC_HBE_USF4 6          /* upsamp_4_flag (must be 0) */
#define XAAC_HBE_XOVER 8         /* x_over_qmf[6] */
#define XAAC_HBE_CFG_WORDS 16
/* d_state [n][XAAC_HBE_ST_WORDS] float, in/out: what the instance carries between calls, in the reference's own order */
#define XAAC_HBE_ST_TAIL 0       /* [32]      ptr_input_buf[no_bins * synth_size ..+ synth_size) */
#define XAAC_HBE_ST_SYNTH 32     /* [384]     synth_buf[0 .. 18 * synth_size) */
#define XAAC_HBE_ST_ANAL 416     /* [384]     analy_buf[0 .. 18 * synth_size) */
#define XAAC_HBE_ST_QIN 800      /* [12][128] qmf_in_buf rows 16..27 (the next call's rows 0..11); only columns
                                  *           4 k_start .. 4 (k_start + synth_size) - 1 are ever non-zero / touched */
#define XAAC_HBE_ST_QOUT 2336    /* [10][128] qmf_out_buf rows 32..41 (the next call's rows 0..9; the rows above are zero) */
#define XAAC_HBE_ST_WORDS 3616
/*   d_qmf_re / d_qmf_im [n][32][64] the frame's 32 new QMF slots (qmf_buf_real/imag + op_delay + SBR_HF_ADJ_OFFSET +
 *                        ESBR_HBE_DELAY_OFFSET, decoder/ixheaacd_sbr_dec.c:899-904)
 *   d_pv_re / d_pv_im   [n][32][64] ph_vocod_qmf_real/imag + op_delay + SBR_HF_ADJ_OFFSET; bands start_band..end_band-1 written
 *   d_err [n] or NULL: 0, -1 / 0x80000000 (the reference's own failure returns), -2 (outside the supported subset) */
int32_t xaac_b200_esbr_hbe_apply_dev(xaac_b200_ctx *ctx, const float *d_qmf_re, const float *d_qmf_im, float *d_pv_re,
                                     float *d_pv_im, const int32_t *d_cfg, float *d_state, int32_t *d_err, int64_t n_units,
                                     void *stream);

/* Whole float eSBR stage: the eSBR branch of ixheaacd_sbr_dec (decoder/ixheaacd_sbr_dec.c:812-1006) for one USAC channel
 * per unit with apply_processing = 1, hbe_flag = 0, no PS / MPS / DRC, stereo_config_idx <= 0, 2:1, 16 time slots:
 *   history shift (memmove of op_delay + SBR_HF_ADJ_OFFSET = 8 rows of qmf_buf_* and sbr_qmf_out_*), 32-band analysis,
 *   ixheaacd_generate_hf, ixheaacd_sbr_env_calc, ixheaacd_esbr_synthesis_regrp, 64-band synthesis, optional ixheaacd_samples_sat.
 * Four kernel launches; the shifts, the regrouping and both hand-overs are fused into the banks' loads / stores.
 * The per-stream state stays in HBM between calls (caller-owned device buffers, all in/out): */
typedef struct xaac_b200_esbr_state_view {
  float *qmf_re, *qmf_im;   /* [n][40][64] ptr_sbr_dec->qmf_buf_real / qmf_buf_imag, rows 0..39 */
  float *out_re, *out_im;   /* [n][40][64] ptr_sbr_dec->sbr_qmf_out_real / sbr_qmf_out_imag, rows 0..39 */
  int32_t *anal_states;     /* [n][320]  str_codec_qmf_bank.anal_filter_states_32 */
  int32_t *anal_pos;        /* [n][2]    as for xaac_b200_esbr_anal32_dev */
  int32_t *synth_states;    /* [n][1280] str_synthesis_qmf_bank.filter_states_32 */
  int32_t *synth_pos;       /* [n][2]    as for xaac_b200_esbr_synth64_dev */
  float *bw_prev;           /* [n][6]    ptr_frame_data->bw_array_prev */
  int32_t *patch;           /* [n][8]    {patch_param.num_patches, start_subband[7]} */
  float *ec_state;          /* [n][640]  e_gain | noise_buf */
} xaac_b200_esbr_state_view;
/*   d_time_in  [n][1024] float core samples, or NULL and d_core_in [n][1024] WORD32 (USAC core output, x 2^-15 in the load)
 *   d_hf_par   [n][XAAC_EHF_PAR_WORDS]; d_ec_ipar [n][XAAC_EEC_IPAR_WORDS] (in/out words updated); d_ec_fpar [n][XAAC_EEC_FPAR_WORDS]
 *   d_rg_par   [n][4] WORD32 {pstr_freq_band_data->qmf_sb_prev, sub_band_start, 2 * border_vec[0], 0}
 *   d_out      [n][2048] float time samples or NULL; d_pcm16 interleaved PCM16 or NULL (ch_fac as for _synth64_pcm16_dev)
 *   d_err      [4][n] or NULL: per unit the err word of the analysis, HF generator, envelope adjuster and synthesis kernels */
int32_t xaac_b200_esbr_dec_dev(xaac_b200_ctx *ctx, const xaac_b200_esbr_state_view *st, const float *d_time_in,
                               const int32_t *d_core_in, const int32_t *d_hf_par, int32_t *d_ec_ipar, const float *d_ec_fpar,
                               const int32_t *d_rg_par, float *d_out, int16_t *d_pcm16, int32_t ch_fac, int32_t *d_err,
                               int64_t n_units, void *stream);

/* The same stage with the harmonic transposer (ptr_header_data->hbe_flag = 1, sbr_patching_mode = 0 frames use it in the HF
 * generator): the core QMF is delayed by ESBR_HBE_DELAY_OFFSET = 32 slots, so base.qmf_re / qmf_im are [n][72][64] (rows 32..71
 * move to rows 0..39 every frame, the analysis bank writes rows 40..71, decoder/ixheaacd_sbr_dec.c:821-846), and the
 * transposer (xaac_b200_esbr_hbe_apply_dev, fused history shift of its output rows) runs between the analysis bank and the HF
 * generator.  Five launches.  d_hf_par must carry XAAC_EHF_HBE_FLAG = 1; d_err is [5][n] (row 4 = the transposer). */
typedef struct xaac_b200_esbr_hbe_state_view {
  xaac_b200_esbr_state_view base;
  float *pv_re, *pv_im;     /* [n][40][64] ptr_sbr_dec->ph_vocod_qmf_real / ph_vocod_qmf_imag, rows 0..39 */
  float *hbe_state;         /* [n][XAAC_HBE_ST_WORDS] */
} xaac_b200_esbr_hbe_state_view;
int32_t xaac_b200_esbr_dec_hbe_dev(xaac_b200_ctx *ctx, const xaac_b200_esbr_hbe_state_view *st, const float *d_time_in,
                                   const int32_t *d_core_in, const int32_t *d_hbe_cfg, const int32_t *d_hf_par,
                                   int32_t *d_ec_ipar, const float *d_ec_fpar, const int32_t *d_rg_par, float *d_out,
                                   int16_t *d_pcm16, int32_t ch_fac, int32_t *d_err, int64_t n_units, void *stream);

/* ---- float parametric stereo of the eSBR branch: ixheaacd_esbr_apply_ps (decoder/ixheaacd_ps_dec_flt.c:381-505) in the 20-band
 * configuration the bitstream parser always selects (use_34_st_bands = 0, use_pca_rot_flg = 0, ixheaacd_sbrdec_lpfuncs.c:636-637),
 * ps_mode = 0: hybrid analysis, transient detection, all-pass decorrelation, rotation, hybrid synthesis, fused with the
 * regrouping (ixheaacd_esbr_synthesis_regrp) and the six look-ahead slots of decoder/ixheaacd_sbr_dec.c:481-517.
 * ROM blob (float / int32 words), built from ia_ps_tables_struct (decoder/ixheaacd_sbr_rom.h:158-239): */
#define XAAC_FPSROM_P8 0         /* p8_13_20[13] */
#define XAAC_FPSROM_P2 16        /* p2_13_20[13] */
#define XAAC_FPSROM_COS2 32      /* cos_mod_2channel[2][13] */
#define XAAC_FPSROM_CS8 64       /* cos_sin_mod_8channel[8][26] */
#define XAAC_FPSROM_QF_RE 272    /* qmf_fract_delay_phase_factor_re[64] */
#define XAAC_FPSROM_QF_IM 336    /* qmf_fract_delay_phase_factor_im[64] */
#define XAAC_FPSROM_SUB_RE 400   /* frac_delay_phase_fac_qmf_sub_re_20[12] */
#define XAAC_FPSROM_SUB_IM 416   /* frac_delay_phase_fac_qmf_sub_im_20[12] */
#define XAAC_FPSROM_QSER_RE 432  /* qmf_ser_fract_delay_phase_factor_re[64][3] */
#define XAAC_FPSROM_QSER_IM 624  /* qmf_ser_fract_delay_phase_factor_im[64][3] */
#define XAAC_FPSROM_SSER_RE 816  /* frac_delay_phase_fac_ser_qmf_sub_re_20[12][3] */
#define XAAC_FPSROM_SSER_IM 856  /* frac_delay_phase_fac_ser_qmf_sub_im_20[12][3] */
#define XAAC_FPSROM_DECAY 896    /* all_pass_link_decay_ser[3] */
#define XAAC_FPSROM_QDELN 900    /* int32 qmf_delay_idx_tbl[64] */
#define XAAC_FPSROM_GRB 964      /* int32 group_borders_20_tbl[23] */
#define XAAC_FPSROM_BGM 988      /* int32 bin_group_map_20[22] (bit 12 = NEGATE_IPD_MASK) */
#define XAAC_FPSROM_DSER 1012    /* int32 delay_sample_ser[3] (rev_link_delay_ser) */
#define XAAC_FPSROM_WORDS 1016
int32_t xaac_b200_set_fps_rom(xaac_b200_ctx *ctx, const void *tables, size_t bytes);
/* d_side [n][XAAC_FPS_SIDE_WORDS]: int32 words num_env, border_position[0..5], usb (sub_band_end); then six sets of
 * [8][20] floats h11r h12r h21r h22r h11i h12i h21i h22i: set 0 = h*_prev of the instance, set 1 + e = the h*_vec of envelope e
 * (ps_dec_flt.c:920-1020; double-precision libm on the host — see libxaac_b200/dropin/ixheaacd_b200_pack_ps_flt.h).
 * Supported: border_position[0] = 0, border_position[num_env] = 32, strictly increasing; anything else -> err -2. */
#define XAAC_FPS_SIDE_NUM_ENV 0
#define XAAC_FPS_SIDE_BORDER 1
#define XAAC_FPS_SIDE_USB 7
#define XAAC_FPS_SIDE_H 16
#define XAAC_FPS_SIDE_WORDS 1024
/* d_state [n][XAAC_FPS_ST_WORDS] in/out (float words unless noted), members of ia_ps_dec_struct: */
#define XAAC_FPS_ST_HYB 0        /* hyb_qmf_buf_re_34[5][12] | _im (bands 0..2 are also hyb_qmf_buf_*_20) */
#define XAAC_FPS_ST_SUBDEL 120   /* sub_qmf_delay_buf_re[2][0..11] | _im */
#define XAAC_FPS_ST_SERSUB 168   /* ser_sub_qmf_dealy_buf_re[3][5][0..11] | _im */
#define XAAC_FPS_ST_QDEL 528     /* qmf_delay_buf_re[14][64] | _im */
#define XAAC_FPS_ST_SERQ 2320    /* ser_qmf_delay_buf_re[3][5][64] | _im */
#define XAAC_FPS_ST_BINS 4240    /* peak_decay_fast_bin[20] | prev_nrg_bin[20] | prev_peak_diff_bin[20] */
#define XAAC_FPS_ST_IDX 4300     /* int32 delay_buf_idx, delay_buf_idx_ser[3], delay_qmf_delay_buf_idx[64] */
#define XAAC_FPS_ST_WORDS 4368
/* The mono channel is read exactly as the synthesis bank's stage mode reads it: slot s, band k = row 2 + s of d_low_* when
 * k < x_over(s), else of d_high_* (d_rg_par as in xaac_b200_esbr_dec_dev); rows 34..39, bands 0..4 of d_low_* are the look-ahead.
 * low_rows = 40 or 72 (rows per unit of d_low_*).  d_left / d_right [n][32][128]: per slot re[64] | im[64]. */
int32_t xaac_b200_esbr_ps_apply_dev(xaac_b200_ctx *ctx, const float *d_low_re, const float *d_low_im, int32_t low_rows,
                                    const float *d_high_re, const float *d_high_im, const int32_t *d_rg_par,
                                    const float *d_side, float *d_state, float *d_left, float *d_right, int32_t *d_err,
                                    int64_t n_units, void *stream);
/* Whole stage for a mono + PS element (channel_mode == PS_STEREO or enh_sbr_ps, decoder/ixheaacd_sbr_dec.c:976-1001): the stage
 * of xaac_b200_esbr_dec_dev / _dec_hbe_dev (st->pv_re = NULL selects the former) up to the envelope adjuster, the PS kernel,
 * then the synthesis bank twice — the left matrix through the element's own bank, the right one through the second channel's.
 * d_err is [6][n] (row 4 = transposer, row 5 = PS). */
typedef struct xaac_b200_esbr_ps_view {
  float *ps_state;          /* [n][XAAC_FPS_ST_WORDS] */
  float *left, *right;      /* [n][32][128] scratch */
  int32_t *synth_states_r;  /* [n][1280] pstr_sbr_channel[1]->str_sbr_dec.str_synthesis_qmf_bank.filter_states_32 */
  int32_t *synth_pos_r;     /* [n][2] */
} xaac_b200_esbr_ps_view;
int32_t xaac_b200_esbr_dec_ps_dev(xaac_b200_ctx *ctx, const xaac_b200_esbr_hbe_state_view *st, const xaac_b200_esbr_ps_view *ps,
                                  const float *d_time_in, const int32_t *d_core_in, const int32_t *d_hbe_cfg,
                                  const int32_t *d_hf_par, int32_t *d_ec_ipar, const float *d_ec_fpar, const int32_t *d_rg_par,
                                  const float *d_ps_side, float *d_out_l, float *d_out_r, int32_t *d_err, int64_t n_units,
                                  void *stream);

/* The stage in two halves.  On reset frames and on frames where sbr_patching_mode changes, ixheaacd_sbr_env_calc first rebuilds
 * its limiter tables from the patch table that ixheaacd_generate_hf of the SAME frame has just produced
 * (ixheaacd_createlimiterbands, decoder/ixheaacd_esbr_envcal.c:169-190, 910-1012: a shell sort and a handful of double-precision
 * log() calls per stream — control plane).  A host that has such units in the batch calls
 *   _front_dev  (analysis bank, [transposer], HF generator; d_err rows 0, 1, 4),
 *   reads st->base.patch of those units, runs ixheaacd_createlimiterbands, stores the tables at XAAC_EEC_LIM_TABLE / _GATE_MODE of
 *   their d_ec_ipar records and sets XAAC_EEC_LIM_REBUILT,
 *   _back_dev   (envelope adjuster, [PS], synthesis bank(s); d_err rows 2, 3, 5).
 * st->pv_re = NULL selects the stage without the transposer, ps = NULL the stage without PS (then d_out_r / d_ps_side are unused
 * and d_pcm16 is allowed).  Calling both halves back to back equals xaac_b200_esbr_dec_dev / _dec_hbe_dev / _dec_ps_dev. */
int32_t xaac_b200_esbr_dec_front_dev(xaac_b200_ctx *ctx, const xaac_b200_esbr_hbe_state_view *st, const float *d_time_in,
                                     const int32_t *d_core_in, const int32_t *d_hbe_cfg, const int32_t *d_hf_par, int32_t *d_err,
                                     int64_t n_units, void *stream);
int32_t xaac_b200_esbr_dec_back_dev(xaac_b200_ctx *ctx, const xaac_b200_esbr_hbe_state_view *st, const xaac_b200_esbr_ps_view *ps,
                                    int32_t *d_ec_ipar, const float *d_ec_fpar, const int32_t *d_rg_par, const float *d_ps_side,
                                    float *d_out, float *d_out_r, int16_t *d_pcm16, int32_t ch_fac, int32_t *d_err,
                                    int64_t n_units, void *stream);

/* apply_processing = 0 (frames before the first SBR header of a stream): the stage only upsamples — history shifts, analysis bank,
 * sbr_qmf_out cleared, synthesis bank over the regrouped core bands (d_rg_par = {x, sub_band_start, 0, 0}); with ps the right
 * channel is the same matrix through the second channel's bank (decoder/ixheaacd_sbr_dec.c:518-525, 836-878, 964-1003).  The
 * transposer, HF generator, envelope adjuster and PS instances are not touched. */
int32_t xaac_b200_esbr_dec_bypass_dev(xaac_b200_ctx *ctx, const xaac_b200_esbr_hbe_state_view *st, const xaac_b200_esbr_ps_view *ps,
                                      const float *d_time_in, const int32_t *d_core_in, const int32_t *d_rg_par, float *d_out,
                                      float *d_out_r, int16_t *d_pcm16, int32_t ch_fac, int32_t *d_err, int64_t n_units,
                                      void *stream);

/* ---- AAC pre-IMDCT spectral stage: ixheaacd_channel_pair_process (decoder/ixheaacd_channel.c:602-718) for AAC-LC
 * elements of 1 or 2 channels, frame length 1024, total_channels <= 2: M/S stereo (ixheaacd_ms_stereo_process,
 * decoder/ixheaacd_stereo.c:54-116), intensity stereo (ixheaacd_intensity_stereo_process, :129-236) and TNS
 * (ixheaacd_aac_tns_process, decoder/ixheaacd_pns_js_thumb.c:248-514, with ixheaacd_tns_decode_coef,
 * ixheaacd_tns_parcor_lpc_convert_dec, ixheaacd_calc_max_spectral_line_dec, ixheaacd_tns_ar_filter_dec) on the dequantised,
 * scale-factor-applied spectrum, in place, and perceptual noise substitution (ixheaacd_map_ms_mask_pns, channel.c:703-726;
 * ixheaacd_pns_process / ixheaacd_gen_rand_vec, pns_js_thumb.c:74-200; ixheaacd_sqrt, decoder/ixheaacd_basic_funcs.c:155-196).
 * Elements that use LTP or the error-resilient syntaxes are not covered; malformed side info is refused with -2, element untouched.
 * ROM: the leading 620 bytes of ia_aac_dec_block_tables_struct (decoder/ixheaacd_aac_rom.h:25-43: pow table, scale_table,
 * tns_max_bands_tbl, tns_coeff3_16, tns_coeff4_16, scale_mant_tab), passed as the reference passes pstr_block_tables.
 * d_side [n][XAAC_SPS_BYTES] per element — the reference's own plain-data members, byte for byte: */
#define XAAC_SPS_NUM_CH 0            /* int32 word 0: channels of the element (1, 2) */
#define XAAC_SPS_COMMON_WINDOW 1     /* int32 word 1: ptr_aac_dec_channel_info[LEFT]->common_window */
#define XAAC_SPS_CORRELATED 16       /* byte offset: pstr_pns_corr_info->correlated[16] of LEFT as the parser left it */
#define XAAC_SPS_MS_USED 32          /* byte offset: pstr_stereo_info->ms_used[8][64] */
#define XAAC_SPS_CH 544              /* byte offset of channel 0's block; channel 1 follows at + XAAC_SPS_CH_BYTES */
#define XAAC_SPS_CH_BYTES 1584
#define XAAC_SPS_CH_WINDOW_SEQUENCE 0 /* int32 words of a channel block: str_ics_info.window_sequence, */
#define XAAC_SPS_CH_MAX_SFB 1         /*   .max_sfb, */
#define XAAC_SPS_CH_NUM_WINDOW_GROUPS 2 /* .num_window_groups, */
#define XAAC_SPS_CH_PNS_ACTIVE 3      /*   str_pns_info.pns_active, */
#define XAAC_SPS_CH_TNS_MAX_BANDS 4   /*   tns_max_bands_tbl[sampling_rate_index][window_sequence is short], */
#define XAAC_SPS_CH_SR_INDEX 5        /*   str_ics_info.sampling_rate_index (informational) */
#define XAAC_SPS_CH_GROUP_LEN 32     /* byte offsets inside a channel block: str_ics_info.window_group_length[8] */
#define XAAC_SPS_CH_CODE_BOOK 40     /*   ptr_code_book[128] (WORD8) */
#define XAAC_SPS_CH_SCALE_FACTOR 168 /*   ptr_scale_factor[128] (WORD16) */
#define XAAC_SPS_CH_TNS 424          /*   ia_tns_info_aac_struct str_tns_info, 924 bytes as laid out by the x86-64 ABI */
#define XAAC_SPS_CH_SFB_INDEX 1348   /*   str_aac_sfb_info[window_sequence].sfb_index[0..51] (WORD16) */
#define XAAC_SPS_CH_PNS_USED 1456    /*   str_pns_info.pns_used[128] */
#define XAAC_SPS_BYTES 3712
int32_t xaac_b200_set_block_rom(xaac_b200_ctx *ctx, const void *block_tables, size_t bytes);
/* d_spec [n][2][1024] WORD32 (ptr_spec_coeff of LEFT, RIGHT; the second half is not touched for single-channel elements), in/out;
 * d_pns_seed [n] in/out: pstr_pns_rand_vec_data->current_seed of the stream the element belongs to (the generator runs on from
 * frame to frame); NULL when no element uses PNS (one that does then gets -2); d_err [n]: 0, or -2.  Three launches (two without
 * d_pns_seed).  pns_frame_number (a plain frame counter) stays with the host. */
int32_t xaac_b200_aac_spectral_dev(xaac_b200_ctx *ctx, int32_t *d_spec, const uint8_t *d_side, int32_t *d_pns_seed, int32_t *d_err,
                                   int64_t n_units, void *stream);

/* ---- SBR side-info dequantisation --------------------------------------------------------------------
 * Batched drop-in for ixheaacd_dec_sbrdata(hdr_ch0, hdr_ch1, frame_ch0, prev_ch0, frame_ch1, prev_ch1, common_tables, ldmps_present,
 * audio_object_type, ec_flag) (decoder/ixheaacd_env_dec.c:628; called from ixheaacd_applysbr, decoder/ixheaacd_sbrdecoder.c:711 and
 * :1263) on the fixed-point path: usac_flag = enh_sbr = 0 (xaacdec -esbr:0), ldmps_present = 0, ec_flag = 0, not AOT_ER_AAC_ELD.
 * Covers ixheaacd_dec_envelope (:727: timing check, ixheaacd_lean_sbrconcealment, ixheaacd_wrong_timing_compensate, coupling-mode
 * change, ixheaacd_process_del_cod_env_data, ixheaacd_check_env_data with its one retry, ixheaacd_dequant_env_data),
 * ixheaacd_calc_noise_floor (:396) and ixheaacd_sbr_env_dequant_coup_fix (:516).  The FLOAT32 shadow arrays of the frame data
 * (flt_env_sf_arr / flt_noise_floor, consumed by the eSBR branch only) are not produced.
 * One record per element, WORD16 words: header, then one block per channel; every field is the reference field of that name.
 * Fields marked io are rewritten in place. */
#define XAAC_SD_NUM_CH 0          /* 1: ptr_sbr_data_ch_1 == NULL, 2: channel pair */
#define XAAC_SD_SHARED_HDR 1      /* 1: ptr_header_data_ch_0 == ptr_header_data_ch_1 (one err_flag / err_flag_prev for both) */
#define XAAC_SD_ERR 2             /* out: what the function returned: 0, 1 = IA_FATAL_ERROR, 2 = -1 (ixheaacd_wrong_timing_compensate) */
#define XAAC_SD_CH 8              /* first channel block */
#define XAAC_SD_CH_WORDS 648
#define XAAC_SD_WORDS (XAAC_SD_CH + 2 * XAAC_SD_CH_WORDS) /* 1304 words = 2608 bytes */
/* channel block */
#define XAAC_SDC_NUM_SF_LO 0      /* pstr_freq_band_data->num_sf_bands[LOW] */
#define XAAC_SDC_NUM_SF_HI 1      /*                      num_sf_bands[HIGH] */
#define XAAC_SDC_NUM_NF 2         /*                      num_nf_bands */
#define XAAC_SDC_NUM_TIME_SLOTS 3 /* header num_time_slots */
#define XAAC_SDC_ERR_FLAG 4       /* io header err_flag */
#define XAAC_SDC_ERR_FLAG_PREV 5  /* io header err_flag_prev */
#define XAAC_SDC_HDR_AMP_RES 6    /* header amp_res */
#define XAAC_SDC_NUM_NOISE_SFAC 7 /* out (channel 0 of a coupled pair): num_noise_sfac */
#define XAAC_SDC_NUM_ENV 8        /* io str_frame_info_details.num_env */
#define XAAC_SDC_NUM_NOISE_ENV 9  /* io                       .num_noise_env */
#define XAAC_SDC_TRANSIENT_ENV 10 /* io                       .transient_env */
#define XAAC_SDC_AMP_RES 11       /* io amp_res */
#define XAAC_SDC_COUPLING 12      /* io coupling_mode */
#define XAAC_SDC_NUM_ENV_SFAC 13  /* io num_env_sfac */
#define XAAC_SDC_MAX_QMF_SB 14    /* io max_qmf_subband_aac */
#define XAAC_SDC_FREQ_RES 16      /* io freq_res[8] */
#define XAAC_SDC_BORDER 24        /* io border_vec[9] */
#define XAAC_SDC_NOISE_BORDER 33  /* io noise_border_vec[3] */
#define XAAC_SDC_DIR 36           /* io del_cod_dir_arr[8] */
#define XAAC_SDC_DIR_NOISE 44     /* io del_cod_dir_noise_arr[2] */
#define XAAC_SDC_INVF 46          /* io sbr_invf_mode[10] */
#define XAAC_SDC_ADD_HARM 56      /* io add_harmonics[56] */
#define XAAC_SDC_ENV 112          /* io int_env_sf_arr[448]: Huffman-decoded deltas in, (mantissa | exponent) words out */
#define XAAC_SDC_NOISE 560        /* io int_noise_floor[10] */
#define XAAC_SDC_PREV_NRG 570     /* io prev sfb_nrg_prev[56] */
#define XAAC_SDC_PREV_NOISE 626   /* io prev prev_noise_level[5] */
#define XAAC_SDC_PREV_AMP_RES 631 /* prev amp_res */
#define XAAC_SDC_PREV_END_POS 632 /* prev end_position */
#define XAAC_SDC_PREV_MAX_QMF 633 /* prev max_qmf_subband_aac */
#define XAAC_SDC_PREV_COUPLING 634 /* prev coupling_mode */
#define XAAC_SDC_PREV_INVF 635    /* prev sbr_invf_mode[10] */
/* d_records [n][XAAC_SD_WORDS] in/out.  Needs xaac_b200_set_env_rom (the misc tables: log_dual_is_table, inv_table).  One launch. */
int32_t xaac_b200_dec_sbrdata_dev(xaac_b200_ctx *ctx, int16_t *d_records, int64_t n_elements, void *stream);

/* PS side info: batched drop-in for ixheaacd_decode_ps_data(ia_ps_dec_struct *, frame_size) (decoder/ixheaacd_ps_bitdec.c:98; called
 * from ixheaacd_applysbr, decoder/ixheaacd_sbrdecoder.c:723): time / frequency delta decoding of the IID and ICC indices with their
 * clamps, 10 -> 20 band expansion, the "no data" and variable-border envelope fix-ups, ixheaacd_map_34_params_to_20
 * (decoder/ixheaacd_sbrdec_lpfuncs.c:561).  One record per PS instance, WORD16 words, every field the ia_ps_dec_struct member of
 * that name (decoder/ixheaacd_ps_dec.h:137-159). */
#define XAAC_PSD_DATA_PRESENT 0   /* io ps_data_present (cleared) */
#define XAAC_PSD_ENABLE_IID 1
#define XAAC_PSD_ENABLE_ICC 2
#define XAAC_PSD_IID_MODE 3       /* 0 / 1 / 2: 10 / 20 / 34 bands */
#define XAAC_PSD_ICC_MODE 4
#define XAAC_PSD_IID_QUANT 5
#define XAAC_PSD_FRAME_CLASS 6
#define XAAC_PSD_NUM_ENV 7        /* io */
#define XAAC_PSD_FRAME_SIZE 8     /* 1024 or 960 (the function's second argument) */
#define XAAC_PSD_BORDER 9         /* io border_position[7] */
#define XAAC_PSD_IID_DT 16        /* iid_dt[5] */
#define XAAC_PSD_ICC_DT 21        /* icc_dt[5] */
#define XAAC_PSD_IID_TABLE 32     /* io iid_par_table[7][34] */
#define XAAC_PSD_ICC_TABLE 270    /* io icc_par_table[7][34] */
#define XAAC_PSD_IID_PREV 508     /* io iid_par_prev[34] */
#define XAAC_PSD_ICC_PREV 542     /* io icc_par_prev[34] */
#define XAAC_PSD_WORDS 576        /* 1152 bytes */
/* d_records [n][XAAC_PSD_WORDS] in/out, 16-byte aligned.  One launch; needs no tables. */
int32_t xaac_b200_decode_ps_data_dev(xaac_b200_ctx *ctx, int16_t *d_records, int64_t n_units, void *stream);

/* ---- raw device-memory helpers for C hosts that do not link the CUDA runtime themselves (the reference-side drop-in glue,
 * libxaac_b200/dropin/ixheaacd_b200_glue.c): allocation and synchronous copies on the context's device ---- */
int32_t xaac_b200_dev_alloc(xaac_b200_ctx *ctx, size_t bytes, void **d_ptr);
int32_t xaac_b200_dev_free(xaac_b200_ctx *ctx, void *d_ptr);
int32_t xaac_b200_h2d(xaac_b200_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int32_t xaac_b200_d2h(xaac_b200_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
int32_t xaac_b200_dev_memset(xaac_b200_ctx *ctx, void *d_ptr, int32_t value, size_t bytes);
/* Peer buffers for multi-GPU hosts (one process per GPU): a buffer from xaac_b200_dev_alloc is exported as a 64-byte CUDA IPC
 * handle; another process imports it on ITS device with peer access enabled, and may then pass the pointer to any _dev entry
 * point: the kernels load / store the peer GPU's HBM over NVLink directly (e.g. pre-parsed per-frame buffers that live on one
 * rank are consumed in place instead of being scattered first). */
int32_t xaac_b200_ipc_export(xaac_b200_ctx *ctx, void *d_ptr, void *handle64);
int32_t xaac_b200_ipc_import(xaac_b200_ctx *ctx, const void *handle64, void **d_ptr);
int32_t xaac_b200_ipc_close(xaac_b200_ctx *ctx, void *d_ptr);

#ifdef __cplusplus
}
#endif
#endif /* XAAC_B200_H */
