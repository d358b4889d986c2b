// kernels.h — internal launch interface between the C-ABI (xaac_b200_api.cu) and the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace xb {

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
};

size_t qmf_synth_table_bytes();
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
};
cudaError_t launch_calc_sbrenvelope_hq(const EnvCalcArgs &args, int num_sms, cudaStream_t stream);

size_t imdct_smem_bytes();
cudaError_t launch_imdct(const ImdctArgs &args, int num_sms, cudaStream_t stream);

}  // namespace xb
