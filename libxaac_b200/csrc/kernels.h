// kernels.h — internal launch interface between the C-ABI (xaac_b200_api.cu) and the kernels.
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>

namespace xb {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: every launcher keeps one of these and opts
// its kernel in once per device the calling thread is on (a process may hold contexts on several GPUs).
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0};
  int dev = 0;
  bool needed() {
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    return ((mask.load(std::memory_order_acquire) >> (dev & 63)) & 1ull) == 0;
  }
  void done() { mask.fetch_or(1ull << (dev & 63), std::memory_order_release); }
};

// Byte offsets inside the IMDCT ROM blob: the leading 7500 bytes of the reference's
// ia_aac_dec_imdct_tables_struct (decoder/ixheaacd_aac_rom.h:112-121), which the host passes in
// exactly as the reference passes `ptr_aac_tables->pstr_imdct_tables` to its own kernels.
constexpr int kRomCos = 0;             // WORD16 cosine_array_2048_256[514]
constexpr int kRomDigRevLong = 1028;   // WORD8  dig_rev_table8_long[64]   (octal digit swap; implied by the kernel)
constexpr int kRomDigRevShort = 1092;  // WORD8  dig_rev_table8_short[8]   (identity)
constexpr int kRomFftTw = 1100;        // WORD32 fft_twiddle[448]
constexpr int kRomWinLongSine = 2892;  // WORD16[1024]
constexpr int kRomWinLongKbd = 4940;   // WORD16[1024]
constexpr int kRomWinShortSine = 6988; // WORD16[128]
constexpr int kRomWinShortKbd = 7244;  // WORD16[128]
constexpr int kRomImdctBytes = 7500;

// Device-side layout of the same tables (re-packed by xaac_b200_set_imdct_rom so that every table is 16-byte
// aligned for vector loads; the digit-reverse tables are not needed on the device).
constexpr int kDevCos = 0;              // 1028 B -> padded to 1040
constexpr int kDevFftTw = 1040;         // 1792 B
constexpr int kDevWinLongSine = 2832;   // 2048 B
constexpr int kDevWinLongKbd = 4880;    // 2048 B
constexpr int kDevWinShortSine = 6928;  // 256 B
constexpr int kDevWinShortKbd = 7184;   // 256 B
constexpr int kDevImdctBytes = 7440;

struct ImdctArgs {
  const int32_t *spec;   // [n_units][1024] spectral coefficients (read-only; the reference destroys them)
  int32_t *overlap;      // [n_units][512]  overlap state, in/out
  uint8_t *wstate;       // [n_units][2]    {window_shape, window_sequence} of the previous frame, in/out
  const uint8_t *ics;    // [n_units][2]    {window_sequence, window_shape} of this frame
  int32_t *out;          // WORD32 time samples, 1024 per unit (layout: see ch_fac)
  int8_t *qshift_adj;    // [n_units]       ia_ics_info_struct.qshift_adj produced by the stage
  const uint8_t *rom;    // device copy of the IMDCT tables in the kDev* layout
  long long n_units;
  int ch_fac;
};

// ---- fixed-point SBR QMF banks -------------------------------------------------------------------------
// Byte offsets inside the QMF ROM blob: the leading 3464 bytes of the reference's ia_qmf_dec_tables_struct
// (decoder/ixheaacd_sbr_rom.h:71-95), handed over by the host like `sbr_tables_ptr->qmf_dec_tables_ptr`.
constexpr int kQRomW32 = 0;             // WORD16 w_32[60]
constexpr int kQRomW16 = 120;           // WORD16 w_16[24]
constexpr int kQRomDigRev2_32 = 168;    // WORD32[4]
constexpr int kQRomDigRev4_16 = 184;    // WORD32[2]
constexpr int kQRomSinCosL64 = 192;     // WORD16 sbr_sin_cos_twiddle_l64[64]
constexpr int kQRomAltSinL64 = 320;     // WORD16 sbr_alt_sin_twiddle_l64[32]
constexpr int kQRomSinCosL32 = 512;     // WORD16 sbr_sin_cos_twiddle_l32[32]
constexpr int kQRomAltSinL32 = 576;     // WORD16 sbr_alt_sin_twiddle_l32[16]
constexpr int kQRomTCosSinL32 = 608;    // WORD16 sbr_t_cos_sin_l32[64]
constexpr int kQRomQmfC = 904;          // WORD16 qmf_c[1280]
constexpr int kQRomBytes = 3464;

struct QmfSynthArgs {
  const int32_t *matrix;   // [n_units][32][128] per slot re[64] | im[64]  (read-only here)
  int16_t *states;         // [n_units][1280]   ia_sbr_qmf_filter_bank_struct.filter_states, in/out
  int16_t *pos;            // [n_units][2]      {ixheaacd_drc_offset, filter_pos_syn - qmf_c}, in/out
  const int16_t *params;   // [n_units][8]      {ov_lb_scale, lb_scale, hb_scale, st_syn_scale, lsb, usb, split, 0}
  int16_t *pcm;            // PCM16, 2048 per unit (ch_fac interleave as for the IMDCT stage)
  const uint8_t *rom;      // device image built by qmf_synth_build_tables()
  long long n_units;
  int ch_fac;
  int fast_bits;           // inputs below 2^fast_bits cannot saturate any add of the modulation
  int zero;                // always 0; an operand the compiler cannot fold (see add3 in qmf_synth_kernel.cu)
  long long mat_stride = 4096;      // words between the matrices of consecutive units (4864 inside the SBR stage)
  long long pcm_unit_stride = 0;    // != 0: unit u writes at pcm + u * pcm_unit_stride with sample stride ch_fac
  const int16_t *gate = nullptr;    // optional: unit u is skipped when gate[u] == 0
  const void *twiddles = nullptr;   // HOST pointer: image built by qmf_synth_build_twiddles(), passed to the kernel by value
};

size_t qmf_synth_table_bytes();
size_t qmf_synth_twiddle_bytes();
void qmf_synth_build_twiddles(const uint8_t *qrom, void *out);
int qmf_synth_build_tables(const uint8_t *qrom, uint8_t *out);  // returns fast_bits (>0) or -1
cudaError_t launch_qmf_synth_hq(const QmfSynthArgs &args, int num_sms, cudaStream_t stream);

struct QmfAnalArgs {
  const int16_t *pcm;      // core-coder time samples, 1024 per unit (ch_fac interleave as for the IMDCT output)
  int16_t *states;         // [n_units][320]  ia_sbr_qmf_filter_bank_struct.anal_filter_states, in/out
  int16_t *pos;            // [n_units][2]    {core_samples_buffer - anal_filter_states, filter_pos - qmf_c}, in/out
  const int16_t *usb;      // [n_units]       qmf_bank->usb (bands rotated by t_cos)
  int32_t *matrix;         // [n_units][32][128]: re at +0..31, im at +64..95 of each slot row
  const uint8_t *rom;      // device image built by qmf_anal_build_tables()
  long long n_units;
  int ch_fac;
  int exact;               // 1: use saturating adds in the modulation (only if the table bound check failed)
  long long mat_stride = 4096;      // words between the matrices of consecutive units
  long long pcm_unit_stride = 0;    // != 0: unit u reads at pcm + u * pcm_unit_stride with sample stride ch_fac
  // != null: the core coder's WORD32 output [n_units][1024] + its qshift_adj [n_units]; the bank converts on load
  // (round16(shl32_sat(x, qshift_adj)), ixheaacd_allocate_sbr_scr, decoder/ixheaacd_api.c:337-370); pcm is ignored, ch_fac = 1
  const int32_t *w32 = nullptr;
  const int8_t *qshift_adj = nullptr;
};

size_t qmf_anal_table_bytes();
int qmf_anal_build_tables(const uint8_t *qrom, uint8_t *out);  // 0: wrapping path exact, 1: need saturating, -1: bad
cudaError_t launch_qmf_anal_hq(const QmfAnalArgs &args, int num_sms, cudaStream_t stream);

struct HfGenArgs {
  const int32_t *lpc;      // [n_units][2][128]  lpc_filt_states_{real,imag}[i] as re[64] | im[64] rows (read-only)
  int32_t *matrix;         // [n_units][38][128] QMF rows (6 overlap + 32 current); high band written in place
  const int16_t *params;   // [n_units][80]      transposer settings + scalar arguments (include/xaac_b200.h)
  int32_t *bw_prev;        // [n_units][6]       ia_sbr_hf_generator_struct.bw_array_prev, in/out
  int16_t *hb_scale;       // [n_units]          sbr_scale_factor->hb_scale set by the stage
  long long n_units;
  int hb_stride = 1;                 // hb_scale[u * hb_stride]
  const int16_t *gate = nullptr;     // optional: unit u is skipped when gate[u * gate_stride] == 0
  long long gate_stride = 0;
};
cudaError_t launch_hf_generator_hq(const HfGenArgs &args, int num_sms, cudaStream_t stream);

// ---- fixed-point HQ envelope adjuster ----------------------------------------------------------------------
// env ROM = the host's ia_env_calc_tables_struct (decoder/ixheaacd_sbr_rom.h:59-68); misc ROM = the leading part of
// ixheaacd_misc_tables up to sqrt_table (decoder/ixheaacd_common_rom.h:27-37).
constexpr int kERomLimGains = 0;      // WORD16[8]
constexpr int kERomSmooth = 24;       // WORD16[4]
constexpr int kERomInvInt = 32;       // WORD16[49]
constexpr int kERomRandPh = 132;      // WORD32[512 + 56]
constexpr int kERomBytes = 2404;
constexpr int kMRomInvTable = 1444;   // WORD16[256]
constexpr int kMRomSqrtTable = 1956;  // WORD16[257]
constexpr int kMRomBytes = 2470;
// ia_sbr_scale_fact_struct as WORD16[8] (include/xaac_b200.h XAAC_SF_*)
constexpr int kSfLb = 0, kSfStLb = 1, kSfOvLb = 2, kSfHb = 3, kSfOvHb = 4, kSfStSyn = 5, kSfPs = 6;
// per-frame SBR side-info record (include/xaac_b200.h XAAC_ENV_*)
constexpr int kEnvNumTimeSlots = 0, kEnvTimeStep = 1, kEnvChannelMode = 2, kEnvLimiterGains = 3, kEnvInterpolFreq = 4,
              kEnvSmoothingMode = 5, kEnvNumSfLo = 6, kEnvNumSfHi = 7, kEnvNumNfBands = 8, kEnvSubBandStart = 9,
              kEnvSubBandEnd = 10, kEnvNumLfBands = 11, kEnvNumEnv = 12, kEnvTransientEnv = 13, kEnvMaxQmfSubband = 14,
              kEnvMaxQmfSubbandPrev = 15, kEnvBorderVec = 16, kEnvFreqRes = 25, kEnvNoiseBorderVec = 33,
              kEnvLimTbl = 36, kEnvFreqLo = 49, kEnvFreqHi = 78, kEnvFreqNoise = 135, kEnvNoiseFloor = 141,
              kEnvAddHarmonics = 151, kEnvSfArr = 207, kEnvPrmWords = 656;
// ia_sbr_calc_env_struct as WORD16[232] (include/xaac_b200.h XAAC_ENV_ST_*)
constexpr int kEnvStFiltMe = 0, kEnvStFiltNoise = 112, kEnvStNoiseE = 168, kEnvStStartUp = 169, kEnvStPhIndex = 170,
              kEnvStTransPrev = 171, kEnvStHarmIndex = 172, kEnvStHarmPrev = 173, kEnvStWords = 232;

struct EnvCalcArgs {
  const int16_t *params;   // [n_units][656]      side info (read-only)
  int16_t *sf;             // [n_units][8]        ia_sbr_scale_fact_struct, hb_scale / ov_hb_scale updated
  int16_t *state;          // [n_units][232]      ia_sbr_calc_env_struct, in/out
  int32_t *matrix;         // [n_units][38][128]  QMF rows; high band adjusted in place
  int32_t *err;            // [n_units] or null   0 / 0x80000000 like the reference's return value
  const uint8_t *env_rom;  // device copy of ia_env_calc_tables_struct
  const uint8_t *misc_rom; // device copy of the leading part of ixheaacd_misc_tables
  long long n_units;
  long long prm_stride = kEnvPrmWords;  // words between the side-info records of consecutive units
  const int16_t *gate = nullptr;        // optional: unit u is skipped when gate[u * gate_stride] == 0
  long long gate_stride = 0;
  const int16_t *max_qmf_prev = nullptr;  // optional override of params[kEnvMaxQmfSubbandPrev]: max_qmf_prev[u * 16]
};
cudaError_t launch_calc_sbrenvelope_hq(const EnvCalcArgs &args, int num_sms, cudaStream_t stream);

// ---- whole fixed-point HQ SBR stage (ixheaacd_sbr_dec) -------------------------------------------------------
// HF generator argument record (include/xaac_b200.h XAAC_HF_*)
constexpr int kHfFactor = 50, kHfNumIfBands = 51, kHfStartIdx = 52, kHfStopIdx = 53, kHfInvf = 54, kHfInvfPrev = 64,
              kHfOvLbScale = 74, kHfLbScale = 75, kHfMaxQmfSubband = 76;
// per-frame side-info record of the stage (include/xaac_b200.h XAAC_SIDE_*)
constexpr int kSideEnv = 0, kSideHf = 656, kSideApply = 736, kSidePs = 737, kSidePsPrm = 744, kSideWords = 1232;
constexpr int kPsPrmIidQuant = 0, kPsPrmNumEnv = 1, kPsPrmBorder = 2, kPsPrmIid = 9, kPsPrmIcc = 247;
// misc words of the channel state (XAAC_SBR_MISC_*)
constexpr int kMiscMaxQmfPrev = 0, kMiscEndPosPrev = 1, kMiscInvfPrev = 2, kMiscCodecUsb = 12, kMiscSynLsb = 13,
              kMiscSynUsb = 14;
constexpr int kSbrMatWords = 38 * 128;  // scratch QMF matrix of one unit: 6 overlap + 32 current slots, re[64] | im[64]
// host blob layout of the channel state (XAAC_SBR_ST_*), WORD16 offsets
constexpr int kSbrStAnalStates = 0, kSbrStAnalPos = 320, kSbrStSynPos = 322, kSbrStSf = 324, kSbrStMisc = 332,
              kSbrStEnv = 348, kSbrStSynStates = 580, kSbrStBwPrev = 1860, kSbrStLpc = 1872, kSbrStOv = 2384,
              kSbrStWords = 3920;
// PS state blob (XAAC_PS_ST_*), WORD16 offsets; the first kPsDspWords live in one device array
constexpr int kPsStAp = 0, kPsStLd = 128, kPsStSd = 464, kPsStSer = 528, kPsStSub = 1488, kPsStSubSer = 1552,
              kPsStHvec = 2032, kPsStIdx = 2320, kPsStPeak = 2332, kPsStHyb = 2452, kPsDspWords = 2596,
              kPsStSynStatesR = 2596, kPsStSynPosR = 3876, kPsStSfR = 3878, kPsStWords = 3888;
constexpr int kPsIdxSer = 0, kPsIdxDelay = 3, kPsIdxDelayLong = 4, kPsIdxScale = 5, kPsIdxUsb = 6, kPsIdxLsbR = 8,
              kPsIdxUsbR = 9;
// PS ROM = leading 1230 bytes of ia_ps_tables_struct (decoder/ixheaacd_sbr_rom.h:177-203), WORD16 offsets
constexpr int kPsRomDecaySf = 0, kPsRomHybResol = 72, kPsRomRevDecay = 75, kPsRomRevDelay = 78, kPsRomBordersGroup = 81,
              kPsRomGroupShift = 104, kPsRomGroupToBin = 110, kPsRomHybToBin = 132, kPsRomDelayToBin = 142,
              kPsRomFracQmf = 174, kPsRomFracSub = 222, kPsRomFracQmfSer = 254, kPsRomFracSubSer = 446, kPsRomScale = 542,
              kPsRomScaleFine = 557, kPsRomAlpha = 588, kPsRomP2_6 = 596, kPsRomP8_13 = 602, kPsRomBytes = 1230;

struct SbrStageArgs {
  const int16_t *side;   // [n][1232] side info
  int32_t *matrix;       // [n][38][128] scratch QMF matrix
  int32_t *ov;           // [n][768]  overlap slots (state)
  int32_t *lpc;          // [n][256]  LPC state rows (state)
  int16_t *sf;           // [n][8]    scale factors (state)
  int16_t *misc;         // [n][16]   misc state words
  int16_t *usb;          // [n]       scratch: analysis bank usb for the analysis kernel
  int16_t *hf_prm;       // [n][80]   scratch: HF generator argument records
  int16_t *synp;         // [n][8]    scratch: synthesis kernel parameter rows
  const int32_t *err;    // [n]       per-unit return value of the envelope adjuster
  long long n_units;
};
cudaError_t launch_sbr_pre(const SbrStageArgs &a, int num_sms, cudaStream_t s);
cudaError_t launch_sbr_scale(const SbrStageArgs &a, int num_sms, cudaStream_t s);
cudaError_t launch_sbr_post(const SbrStageArgs &a, int num_sms, cudaStream_t s);
// sbr_pre + analysis bank + sbr_scale of one unit by one warp (qmf_anal_kernel.cu); a.usb is not used
cudaError_t launch_sbr_front_hq(const QmfAnalArgs &a, const SbrStageArgs &g, int num_sms, cudaStream_t s);
// envelope adjuster + sbr_post of one unit by one warp (envcalc_kernel.cu); a.err receives the adjuster's return value
cudaError_t launch_calc_sbrenvelope_hq_post(const EnvCalcArgs &a, const SbrStageArgs &g, int num_sms, cudaStream_t s);
cudaError_t launch_pcm16_from_imdct(const int32_t *in, const int8_t *qshift_adj, int16_t *out, long long n_units, int mode,
                                    int num_sms, cudaStream_t s);

struct PsArgs {
  const int16_t *side;     // [n][1232]
  int32_t *matrix;         // [n][38][128] left matrix, rows 0..31 rewritten in the synthesis scale
  int32_t *right;          // [n][32][128] right matrix (out)
  int16_t *ps_state;       // [n][2596]    PS state, in/out
  int16_t *sf;             // [n][8]       left scale factors (ps_scale written)
  int16_t *sf_r;           // [n][8]       right channel's scale factors
  int16_t *synp;           // [n][8]       in: {ov_lb, lb, hb, st_syn, lsb, usb, 6, 0}; out: left synthesis parameters
  int16_t *synp_r;         // [n][8]       out: right synthesis parameters
  int16_t *ps_done;        // [n]          out: 1 when the unit ran PS
  const int32_t *err;
  const uint8_t *ps_rom, *env_rom, *misc_rom;
  long long n_units;
  int rot_nosat = 0;       // 1: no fractional-delay phase factor of the ROM is -32768 (checked by xaac_b200_set_ps_rom)
};
cudaError_t launch_ps_frame(const PsArgs &args, int num_sms, cudaStream_t stream);

// ---- whole fixed-point LOW-POWER SBR stage (ixheaacd_sbr_dec, low_pow_flag = 1), one fused kernel ----------------
struct SbrLpArgs {
  const int16_t *side;       // [n][1232] side info (XAAC_SIDE_*; the PS part is ignored)
  const int16_t *time_in;    // core-coder PCM16: unit u reads 1024 samples at time_in + u * in_unit_stride, stride in_ch
  int16_t *time_out;         // PCM16 out: unit u writes 2048 samples at time_out + (u / out_ch) * 2048 * out_ch + u % out_ch,
                             // sample stride out_ch (out_ch = 2: L/R interleaved stereo frames)
  int16_t *anal_states, *anal_pos, *syn_pos, *sf, *misc, *env, *syn_states;  // channel state, structure of arrays
  int32_t *bw_prev, *lpc, *ov;
  int32_t *err;              // [n] or null: 0 / 0x80000000
  const uint8_t *lp_rom;     // device image built by sbr_lp_build_tables()
  const uint8_t *env_rom, *misc_rom;
  long long n_units;
  long long in_unit_stride = 1024;
  int in_ch = 1;
  int out_ch = 1;
  // != null: the core coder's WORD32 output [n][1024] + qshift_adj [n] instead of time_in; converted on load
  // (round16(shl32_sat(x, qshift_adj)), ixheaacd_allocate_sbr_scr, decoder/ixheaacd_api.c:337-370)
  const int32_t *w32 = nullptr;
  const int8_t *qshift_adj = nullptr;
};
size_t sbr_lp_table_bytes();
int sbr_lp_build_tables(const uint8_t *qrom, uint8_t *out);  // 0 ok, -1 tables unsupported
cudaError_t launch_sbr_dec_lp(const SbrLpArgs &args, int num_sms, cudaStream_t stream);

// ---- USAC frequency-domain core transform (ixheaacd_fd_frm_dec) ------------------------------------------------------
// ROM blob = the reference's const tables of the path in this order (include/xaac_b200.h XAAC_UROM_*)
constexpr int kURomFftTw = 0;        // WORD32[514]  ixheaacd_twiddle_table_fft_32x32
constexpr int kURomCos512 = 2056;    // WORD32[512]  ixheaacd_pre_post_twid_cos_512
constexpr int kURomSin512 = 4104;    // WORD32[512]  ixheaacd_pre_post_twid_sin_512
constexpr int kURomCos64 = 6152;     // WORD32[64]   ixheaacd_pre_post_twid_cos_64
constexpr int kURomSin64 = 6408;     // WORD32[64]   ixheaacd_pre_post_twid_sin_64
constexpr int kURomSine1024 = 6664;  // WORD32[1024] ixheaacd_sine_win_1024
constexpr int kURomKbd1024 = 10760;  // WORD32[1024] ixheaacd_kbd_win1024
constexpr int kURomSine128 = 14856;  // WORD32[128]  ixheaacd_sine_win_128
constexpr int kURomKbd128 = 15368;   // WORD32[128]  ixheaacd_kbd_win128
constexpr int kURomBytes = 15880;
struct UsacFdArgs {
  const int32_t *coef;   // [n][1024] dequantised spectrum usac_data->coef_fix[ch] (read-only here; the reference destroys it)
  int32_t *overlap;      // [n][1024] usac_data->overlap_data_ptr[ch], in/out
  uint8_t *wstate;       // [n]       usac_data->window_shape_prev[ch], in/out
  const uint8_t *ics;    // [n][2]    {window_sequence, window_shape} of this frame
  int32_t *out;          // [n][1024] usac_data->output_data_ptr[ch] (WORD32, Q15)
  const uint8_t *rom;    // device copy of the ROM blob
  long long n_units;
};
size_t usac_fd_smem_bytes();
int usac_fd_check_tables(const uint8_t *urom);
cudaError_t launch_usac_fd(const UsacFdArgs &args, int num_sms, cudaStream_t stream);

// ---- AAC-LC output stage: peak limiter + round16 -----------------------------------------------------------------------
// per-stream state record = ia_peak_limiter_struct as 32-bit words (include/xaac_b200.h XAAC_PL_*)
constexpr int kPlAttackConst = 0, kPlReleaseConst = 1, kPlGainMod = 2, kPlMinGain = 3, kPlPsg = 4, kPlAttack = 6,
              kPlDelayIdx = 7, kPlMaxIdx = 8, kPlCir = 9, kPlLimiterOn = 10, kPlNumCh = 11, kPlMaxBuf = 12,
              kPlMaxAttack = 512, kPlDelayed = 524, kPlWords = 1548;
struct PeakLimArgs {
  int32_t *state;            // [n][1548] in/out
  const int32_t *samples;    // [n][1024][ch] WORD32 time samples (interleaved like the reference's time_data), read-only
  const int8_t *qshift_adj;  // [n][ch]
  int32_t *out32;            // [n][1024][ch] limited WORD32 samples or null
  int16_t *pcm16;            // [n][1024][ch] round16 of them or null
  int32_t *err;              // [n] or null
  long long n_units;
  int ch;
  // deferred smoothing (set by launch_peak_limiter from the scratch block; null: everything inside the main kernel)
  float *gbuf = nullptr;     // [n][1024] raw, then smoothed gains of the deferred streams
  int *list = nullptr;       // [n] queued stream indices
  int *count = nullptr;      // queue length
};
size_t peak_limiter_scratch_bytes(long long n_units);
cudaError_t launch_peak_limiter(const PeakLimArgs &args, void *scratch, int which, int num_sms, cudaStream_t stream);

// ---- SBR side-info dequantisation (sbr_sideinfo_kernel.cu): records [n][XAAC_SD_WORDS], in/out ----
cudaError_t launch_sbr_sideinfo(int16_t *records, long long n, const uint8_t *misc_rom, int num_sms, cudaStream_t stream);
cudaError_t launch_ps_sideinfo(int16_t *records, long long n, int num_sms, cudaStream_t stream);  // [n][XAAC_PSD_WORDS]

// ---- eSBR 64-band synthesis bank (per-slot core of ixheaacd_esbr_synthesis_filt_block) --------------------------------
// ROM blob: esbr_qmf_c[1280] | esbr_w_32[60] | esbr_sin_cos_twiddle_l64[64] | esbr_alt_sin_twiddle_l64[32] | esbr_w_16[24] |
// esbr_sin_cos_twiddle_l32[32] | esbr_alt_sin_twiddle_l32[16] | esbr_t_cos_sin_l32[64] (WORD32)
constexpr int kEsRomQmfC = 0, kEsRomW32 = 5120, kEsRomSinCos = 5360, kEsRomAlt = 5616, kEsRomW16 = 5744, kEsRomSinCos32 = 5840,
              kEsRomAlt32 = 5968, kEsRomTCos32 = 6032, kEsRomBytes = 6288;
struct EsbrSynthArgs {
  const float *qmf;     // [n][32][128] per slot re[64] | im[64] (qmf_buf_real[i][k], qmf_buf_imag[i][k])
  int32_t *states;      // [n][1280] filter_states_32, in/out
  int32_t *pos;         // [n][2] {ixheaacd_drc_offset, filter_pos_syn_32 - esbr_qmf_c}, in/out
  float *out;           // [n][2048] time samples, or null when only PCM is wanted
  int32_t *err;         // [n] or null
  const uint8_t *rom;   // device image built by esbr_synth_build_tables()
  long long n_units;
  int periodic;
  // stage mode (ixheaacd_esbr_synthesis_regrp fused into the load, sbr_dec.c:297-397, stereo_config_idx <= 0): slot s reads row
  // 2 + s, band k from rg_low_* (qmf_buf) when k < x_over(s), else from rg_high_* (sbr_qmf_out); rg_par[u] = {x_over of the
  // slots before stop_border (qmf_sb_prev), x_over of the others (sub_band_start), stop_border (2 * border_vec[0]), 0}
  const float *rg_low_re = nullptr, *rg_low_im = nullptr, *rg_high_re = nullptr, *rg_high_im = nullptr;  // [n][40][64]
  const int32_t *rg_par = nullptr;                                                                       // [n][4]
  long long rg_low_stride = 2560;  // floats per unit of rg_low_* (72 x 64 with the harmonic transposer)
  int16_t *pcm16 = nullptr;  // optional fused ixheaacd_samples_sat: unit u = stream u / pcm_ch_fac, channel u % pcm_ch_fac;
  int pcm_ch_fac = 1;        // sample i of the unit goes to pcm16[(stream * 2048 + i) * pcm_ch_fac + channel]
};
size_t esbr_synth_table_bytes();
int esbr_synth_build_tables(const uint8_t *erom, uint8_t *out);
cudaError_t launch_esbr_synth(const EsbrSynthArgs &args, int num_sms, cudaStream_t stream);
struct EsbrAnalArgs {
  const float *time_in;  // [n][1024] core-coder samples (ptr_sbr_dec->time_sample_buf); or one of the fused hand-overs:
  const int32_t *core_in = nullptr;  // [n][1024] WORD32 USAC core output, x 2^-15 (ixheaacd_ext_ch_ele.c:1040-1046)
  const int16_t *pcm_in = nullptr;   // legacy core PCM16, interleaved: unit u = stream u / pcm_ch_fac, channel u % pcm_ch_fac
  int pcm_ch_fac = 1;                //   (FLOAT32)time_data[ch_fac * i + ch], ixheaacd_api.c:3384-3437
  // stage mode: the unit's qmf_buf_real / imag arrays [n][40][64]; rows 32..39 move to rows 0..7 first (the memmove at the
  // top of ixheaacd_sbr_dec's eSBR branch, sbr_dec.c:836-846, op_delay 6 + SBR_HF_ADJ_OFFSET 2), slot s goes to row 8 + s
  float *stage_re = nullptr, *stage_im = nullptr;
  // with the harmonic transposer the arrays are [n][72][64]: the core QMF is delayed by ESBR_HBE_DELAY_OFFSET = 32 slots
  // (codec_x_delay, sbr_dec.c:821-823), rows 32..71 move to 0..39 and slot s goes to row 40 + s
  int stage_hist_rows = 8;           // 8 or 40
  int32_t *states;       // [n][320] anal_filter_states_32, in/out
  int32_t *pos;          // [n][2] {state_new_samples_pos_low_32 - anal_filter_states_32, filter_pos_32 - esbr_qmf_c}, in/out
  float *qmf;            // unit u writes slot s at qmf + u * out_stride + 128 * s: re at +0..31, im at +64..95
  int32_t *err;          // [n] or null
  const uint8_t *rom;
  long long n_units;
  long long out_stride = 4096;
  int periodic;
};
cudaError_t launch_esbr_anal(const EsbrAnalArgs &args, int num_sms, cudaStream_t stream);

// eSBR float HF generator (ixheaacd_generate_hf): par[] word offsets = XAAC_EHF_* of include/xaac_b200.h
constexpr int kEhfNumMf = 0, kEhfNumIf = 1, kEhfSbStart = 2, kEhfBorderFirst = 3, kEhfBorderLast = 4, kEhfHbeFlag = 5,
              kEhfPatchingMode = 6, kEhfFs = 7, kEhfPreProc = 8, kEhfUsf4 = 9, kEhfMpsSbr = 10, kEhfCovCount = 11,
              kEhfInvf = 16, kEhfInvfPrev = 21, kEhfInvfTbl = 26, kEhfFmaster = 32, kEhfParWords = 96;
struct EsbrHfgenArgs {
  const float *src_re, *src_im;  // [n][40][64] low-band QMF (qmf_buf_real / imag from their first row)
  const float *pv_re, *pv_im;    // [n][40][64] phase-vocoder QMF (ph_vocod_qmf_real / imag) or null
  float *dst_re, *dst_im;        // [n][40][64] sbr_qmf_out_real / imag, in/out (only the cells the reference writes)
  const int32_t *par;            // [n][96]
  float *bw_prev;                // [n][6] bw_array_prev, in/out
  int32_t *patch_out;            // [n][8] {num_patches, start_subband[7]} or null
  int32_t *err;                  // [n] or null
  long long n_units;
  int shift_rows = 0;            // stage mode: rows 32..39 of dst move to rows 0..7 first (sbr_dec.c:848-856)
  long long src_stride = 2560;   // floats per unit of src_* (72 x 64 with the harmonic transposer's delayed core QMF)
};
cudaError_t launch_esbr_hfgen(const EsbrHfgenArgs &args, int num_sms, cudaStream_t stream);

// eSBR float envelope adjuster (ixheaacd_sbr_env_calc, ORIG_SBR): word offsets = XAAC_EEC_* of include/xaac_b200.h
constexpr int kEecSbStart = 0, kEecSbEnd = 1, kEecNumEnv = 2, kEecTransEnv = 3, kEecShortPrev = 4, kEecNumNoiseEnv = 5,
              kEecNumSfLo = 6, kEecNumSfHi = 7, kEecNumNf = 8, kEecSmoothingMode = 9, kEecInterpolFreq = 10,
              kEecLimiterBands = 11, kEecLimiterGains = 12, kEecHarmIndex = 13, kEecPhaseIndex = 14, kEecStartUp = 15,
              kEecReset = 16, kEecSbrMode = 17, kEecUsf4 = 18, kEecPatchingChanged = 19, kEecLimRebuilt = 20, kEecBorder = 24, kEecFreqRes = 33,
              kEecNoiseBorder = 41, kEecInterTes = 44, kEecGateMode = 52, kEecLimTable = 56, kEecTblNoise = 108,
              kEecTblLo = 116, kEecTblHi = 148, kEecAddHarm = 208, kEecHarmPrev = 264, kEecIparWords = 288,
              kEecSfbNrg = 0, kEecNoiseFloor = 448, kEecFparWords = 464, kEecStateWords = 640, kEecRphaseBytes = 4096;
struct EsbrEnvcalcArgs {
  float *re, *im;       // [n][40][64] sbr_qmf_out_real / imag from their first row, in/out
  int32_t *ipar;        // [n][288], in/out words: env_short_flag_prev, harm_index, phase_index, esbr_start_up, harm_flag_prev
  const float *fpar;    // [n][464] flt_env_sf_arr | flt_noise_floor
  float *state;         // [n][640] e_gain[5][64] | noise_buf[5][64], in/out
  const float *rphase;  // ixheaac_random_phase[512][2]
  int32_t *err;         // [n] or null
  long long n_units;
  // inter-TES (envelopes with inter_temp_shape_mode != 0): the low band, qmf_buf_real / imag from their first row; unit stride in
  // floats (2560, or 4608 with the harmonic transposer's delayed core QMF).  Null: such frames return -2.
  const float *low_re = nullptr, *low_im = nullptr;
  long long low_stride = 2560;
};
cudaError_t launch_esbr_envcalc(const EsbrEnvcalcArgs &args, int num_sms, cudaStream_t stream);

size_t imdct_smem_bytes();
cudaError_t launch_imdct(const ImdctArgs &args, int num_sms, cudaStream_t stream);

// QMF harmonic transposer (ixheaacd_qmf_hbe_apply): ROM word offsets = XAAC_HROM_*, cfg words = XAAC_HBE_*, state words =
// XAAC_HBE_ST_* of include/xaac_b200.h
constexpr int kHromWin = 0, kHromSynCos = 1560, kHromAnaCs = 1720, kHromCosTrans = 2040, kHromFftTw = 2488, kHromTw24 = 3004,
              kHromTw48 = 3036, kHromPvCos = 3100, kHromPvSin = 3164, kHromInterp = 3228, kHromSelCase = 3236, kHromXp2 = 3276,
              kHromXp3 = 3788, kHromXp4 = 4300, kHromXp41 = 4812, kHromSyn20 = 5324, kHromAna40 = 6124, kHromWords = 9324;
constexpr int kHbeSynthSize = 0, kHbeKStart = 1, kHbeStartBand = 2, kHbeEndBand = 3, kHbeMaxStretch = 4, kHbePitch = 5,
              kHbeUsf4 = 6, kHbeXover = 8, kHbeCfgWords = 16;
constexpr int kHbeStTail = 0, kHbeStSynth = 32, kHbeStAnal = 416, kHbeStQin = 800, kHbeStQout = 2336, kHbeStWords = 3616;
struct EsbrHbeArgs {
  const float *qmf_re, *qmf_im;  // first of the 32 new QMF rows of unit 0; unit u at + u * in_stride floats, rows of 64
  float *pv_re, *pv_im;          // first of the 32 output rows of unit 0; unit u at + u * out_stride floats
  long long in_stride, out_stride;
  const int32_t *cfg;            // [n][16]
  float *state;                  // [n][3616] in/out
  int32_t *err;                  // [n] or null
  const float *rom;              // device copy of the XAAC_HROM_* blob
  long long n_units;
  // stage mode: pv_* point at row 8 of [n][40][64] arrays (ph_vocod_qmf_real / imag); rows 32..39 move to rows 0..7 first
  // (sbr_dec.c:858-867)
  int shift_rows = 0;
};
cudaError_t launch_esbr_hbe(const EsbrHbeArgs &args, int num_sms, cudaStream_t stream);

// eSBR float parametric stereo (ixheaacd_esbr_apply_ps, 20-band configuration): ROM word offsets = XAAC_FPSROM_*, side words =
// XAAC_FPS_SIDE_*, state words = XAAC_FPS_ST_* of include/xaac_b200.h
constexpr int kFpsRomP8 = 0, kFpsRomP2 = 16, kFpsRomCos2 = 32, kFpsRomCs8 = 64, kFpsRomQfRe = 272, kFpsRomQfIm = 336,
              kFpsRomSubRe = 400, kFpsRomSubIm = 416, kFpsRomQSerRe = 432, kFpsRomQSerIm = 624, kFpsRomSSerRe = 816,
              kFpsRomSSerIm = 856, kFpsRomDecay = 896, kFpsRomQdelN = 900, kFpsRomGrb = 964, kFpsRomBgm = 988,
              kFpsRomDser = 1012, kFpsRomWords = 1016;
constexpr int kFpsSideNumEnv = 0, kFpsSideBorder = 1, kFpsSideUsb = 7, kFpsSideH = 16, kFpsSideWords = 1024;
constexpr int kFpsStHyb = 0, kFpsStSubDel = 120, kFpsStSerSub = 168, kFpsStQDel = 528, kFpsStSerQ = 2320, kFpsStBins = 4240,
              kFpsStIdx = 4300, kFpsStWords = 4368;
struct EsbrPsArgs {
  // the mono channel before regrouping, exactly what the synthesis bank's stage mode reads (EsbrSynthArgs::rg_*): slot s, band k
  // comes from row 2 + s of low_* when k < x_over(s), else of high_*; rows 34..39 of low_* (bands 0..4) are the look-ahead
  const float *low_re, *low_im;    // [n][>= 40][64], unit stride low_stride
  const float *high_re, *high_im;  // [n][40][64]
  const int32_t *rg_par;           // [n][4]
  long long low_stride = 2560;
  const float *side;               // [n][1024]: num_env, border_position[0..5], usb (int32 words) | previous + per-envelope h11..h22
  float *state;                    // [n][4368] in/out
  float *left, *right;             // [n][32][128] per slot re[64] | im[64]: the synthesis bank kernel's input layout
  int32_t *err;                    // [n] or null
  const float *rom;                // device copy of the XAAC_FPSROM_* blob
  long long n_units;
};
cudaError_t launch_esbr_ps(const EsbrPsArgs &args, int num_sms, cudaStream_t stream);

// AAC pre-IMDCT spectral stage (ixheaacd_channel_pair_process for AAC-LC): record byte offsets = XAAC_SPS_* of include/xaac_b200.h,
// ROM = the leading 620 bytes of ia_aac_dec_block_tables_struct (decoder/ixheaacd_aac_rom.h:25-43)
constexpr int kSpsNumCh = 0, kSpsCommonWindow = 1;                       // int32 words of the header
constexpr int kSpsCorrelated = 16, kSpsMsUsed = 32, kSpsCh = 544, kSpsChBytes = 1584, kSpsBytes = 544 + 2 * 1584;
constexpr int kSpsChGroupLen = 32, kSpsChCodeBook = 40, kSpsChScaleFactor = 168, kSpsChTns = 424, kSpsChSfbIndex = 1348,
              kSpsChPnsUsed = 1456;
constexpr int kBromBytes = 620, kBromScaleTable = 129 /* int32 index */, kBromTnsCoeff3 = 278, kBromTnsCoeff4 = 286 /* int16 index */,
              kBromScaleMant = 151 /* int32 index */;
struct AacSpectralArgs {
  int32_t *spec;              // [n][2][1024] ptr_spec_coeff of the element's channels (second unused for a single channel), in/out
  const unsigned char *side;  // [n][kSpsBytes]
  int32_t *err;               // [n]: 0, -2 (outside the supported subset: nothing touched)
  int32_t *pns_seed;          // [n] pstr_pns_rand_vec_data->current_seed, in/out; null: elements that use PNS get -2
  const int32_t *rom;
  long long n_units;
};
cudaError_t launch_aac_spectral(const AacSpectralArgs &args, int num_sms, cudaStream_t stream);

}  // namespace xb
