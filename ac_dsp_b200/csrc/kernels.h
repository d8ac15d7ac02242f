// kernels.h -- launch entry points of the CUDA kernels (host runtime <-> kernel translation units).
#pragma once
#include "common.cuh"

namespace b2d {

// One run() of a FIR handle on device-resident buffers.
struct FirLaunch {
  Fmt fin, fcoeff, facc, fout;
  int n_taps, ftype;
  int ascending;         // ac_fir_reg_share walks the taps upwards (matters for order-dependent accumulators only)
  uint32_t C;            // channels
  int interleaved;       // layout of in / out
  const void *in;        // n samples per channel, input container
  void *out;             // n samples per channel, output container
  size_t n;
  const void *tail;      // [C][n_taps-1] previous samples, input container, planar, oldest first
  void *tail_next;       // same shape: history after this call
  const int64_t *coeff64;   // [C][n_taps] raw coefficients (generic path)
  const uint32_t *coeff_pk; // [C][pk_words] byte-plane packed, reversed coefficients (q15 path), or null
  int pk_words;
  const int32_t *coeff32;   // [C][fir_wide_words] taps of the wide path (fir_wide_pack), or null
};

// generic path: every format / ftype / Q / O, reference tap order, 128-bit intermediates.
cudaError_t launch_fir_generic(const FirLaunch &p, cudaStream_t st);
// TRANSPOSED across a coefficient change (fir_generic.cu): outputs i < n_limit with accumulators started from
// acc_init[C][n_taps-1] (may be null); acc_out != null: zero inputs, raw ACC_TYPE accumulators to acc_out[C][n_taps-1].
cudaError_t launch_fir_pending(const FirLaunch &p, size_t n_limit, const int64_t *acc_init, int64_t *acc_out, cudaStream_t st);
cudaError_t launch_fir_pending_shift(const int64_t *src, int64_t *dst, size_t n, int T, uint32_t C, cudaStream_t st);
// q15 path: W_in, W_c <= 16 in int16 containers, exact left-shift accumulate, DP2A byte planes.
bool fir_q15_supported(const Fmt &in, const Fmt &coeff, const Fmt &acc, const Fmt &out, int n_taps, int ftype);
void fir_q15_pack(const Fmt &coeff, const int64_t *c, int n_taps, int ftype, uint32_t *pk, int pk_words);
int fir_q15_pk_words(int n_taps, int ftype);
cudaError_t launch_fir_q15(const FirLaunch &p, cudaStream_t st);
// overlap-save path (fir_ovs.cu): q15 formats, long filters, FP64 FFT blocks of 4096 with an a-priori error bound < 1/2
// evaluated on the loaded coefficients; tables and spectra are prepared by the host runtime at load time.
void fir_effective_taps(const int64_t *c, int n_taps, int ftype, int64_t *eff);
bool fir_ovs_geometry(int n_taps, uint32_t C, int interleaved);
int fir_ovs_discard(int n_taps);
double fir_ovs_error_bound(const Fmt &in, double l1);
void fir_ovs_tables(double2 *tw1 /*[6][256]*/, double2 *tw2 /*[6][16]*/);
void fir_ovs_spectrum(const int64_t *eff, int n_taps, double2 *hs /*[16][256]*/);
cudaError_t launch_fir_ovs(const FirLaunch &p, const double2 *tw /*[6][256] + [6][16]*/, const double2 *hs, double *resid, cudaStream_t st);
// q24 path: W_in 17..24 in int32 containers, W_c <= 16: coefficient pairs in the DP2A 16-bit lanes, three sample byte planes.
bool fir_q24_supported(const Fmt &in, const Fmt &coeff, const Fmt &acc, const Fmt &out, int n_taps, int ftype);
void fir_q24_pack(const int64_t *c, int n_taps, int ftype, uint32_t *pk, int pk_words);
int fir_q24_pk_words(int n_taps);
cudaError_t launch_fir_q24(const FirLaunch &p, cudaStream_t st);
// wide path: operands <= 32 bits, wrapping 64-bit accumulator with Q in {TRN, RND}; IMAD.WIDE per tap.
bool fir_wide_supported(const Fmt &in, const Fmt &coeff, const Fmt &acc, const Fmt &out, int n_taps, int ftype);
int fir_wide_mode(const Fmt &in, const Fmt &coeff, const Fmt &acc, int n_taps, int ftype);
int fir_wide_words(int n_taps);
void fir_wide_pack(const int64_t *c, int n_taps, int ftype, int mode, int32_t *out, int words);
cudaError_t launch_fir_wide(const FirLaunch &p, cudaStream_t st);
// history carry: tail_next = last (n_taps-1) samples of (tail ++ in).
cudaError_t launch_fir_tail(const FirLaunch &p, cudaStream_t st);
// ac_firProgCoeffs_delay_line: dl[c] = OUT_TYPE(sample N_TAPS-1 steps back) after this call (before the tail swap)
cudaError_t launch_fir_delay_out(const FirLaunch &p, int64_t *dl, cudaStream_t st);

struct CicLaunch {
  Fmt fin, fout;
  int intW;              // lossless internal width (find_inter_type_cic_*)
  int R, M, N, intr;
  uint32_t C;
  int interleaved;
  const void *in;        // n inputs per channel
  void *out;             // n_out outputs per channel (planar stride n_out)
  size_t n, n_out;
  unsigned long long n_seen;   // inputs consumed per channel before this call
  unsigned long long out_first; // global index of the first output of this call
  const void *tail;      // [C][H] previous inputs, planar, oldest first
  void *tail_next;
  int H;                 // history length kept (inputs)
};

int cic_history_len(int intr, int R, int M, int N);
cudaError_t launch_cic_generic(const CicLaunch &p, cudaStream_t st);
bool cic_fast_supported(const CicLaunch &p);
cudaError_t launch_cic_fast(const CicLaunch &p, cudaStream_t st);
bool cic_intr_fast_supported(const CicLaunch &p);
cudaError_t launch_cic_intr_fast(const CicLaunch &p, cudaStream_t st);
cudaError_t launch_cic_tail(const CicLaunch &p, cudaStream_t st);

// Polyphase decimating FIR (ac_poly_dec): fir_dec.cu
struct DecLaunch {
  Fmt fin, fcoeff, facc, fout;
  int nt, df, wide;      // taps per phase, decimation factor, 2 = DP2A kernel / 1 = IMAD.WIDE kernel / 0 = generic kernel
  uint32_t C;
  int interleaved;
  const void *in;        // n inputs per channel
  void *out;             // n_out outputs per channel, planar stride n_out
  size_t n, n_out;
  unsigned long long n_seen;
  const void *tail;      // [C][nt*df - 1] previous samples
  const int64_t *coeff64;   // [C][nt*df] raw taps in the reference's phase order
  const int32_t *coeff32;   // [C][df][polydec_words(nt)] or null
  const uint32_t *coeff_pk; // [C][polydec_q15_words(nt, df)] or null
};
bool polydec_q15_supported(const Fmt &in, const Fmt &coeff, const Fmt &acc, int ntaps, int df);
int polydec_q15_words(int ntaps, int df);
void polydec_q15_pack(const Fmt &coeff, const int64_t *c, int ntaps, int df, uint32_t *out);
int polydec_wide_mode(const Fmt &in, const Fmt &coeff, const Fmt &acc, int ntaps, int df);
int polydec_words(int ntaps);
void polydec_pack(const int64_t *c, int ntaps, int df, int32_t *out);
cudaError_t launch_polydec(const DecLaunch &p, cudaStream_t st);

// Polyphase interpolating FIR (ac_poly_intr): fir_intr.cu
struct PiLaunch {
  Fmt fin, fcoeff, facc, fout;
  int nt, ifac, ftype, csz, fast;   // taps per phase, interpolation factor, B2D_PI_*, coefficients per channel, 64-bit path
  uint32_t C;
  int interleaved;
  const void *in;            // n inputs per channel
  void *out;                 // n_rows * ifac outputs per channel, planar
  size_t n, n_rows;
  int row_shift;             // source step of output row r is r - row_shift
  const void *tail;          // [C][H] previous inputs
  int H;
  const int64_t *coeff64;    // [C][csz]
  const uint8_t *sign, *corr;   // [C][ifac]
  const int64_t *carry;      // [C][ifac] accumulators of the step before this call
  int64_t *carry_next;
};
bool polyintr_fast_supported(const Fmt &in, const Fmt &coeff, const Fmt &acc, int ftype);
cudaError_t launch_polyintr(const PiLaunch &p, cudaStream_t st);

// Integrate-and-dump (ac_intg_dump): intg_dump.cu
struct IdLaunch {
  Fmt fin, facc, fout;
  int chn, force_thread;
  const void *in;                    // samples of this call, interleaved over chn
  void *out;                         // [nseg_out][chn]
  const int64_t *carry;              // [chn] ACC raw on entry
  int64_t *carry_next;               // [chn] ACC raw on exit
  const unsigned long long *table;   // device: [nseg + 1] boundaries, or null when every dumping segment has n_reg samples
  unsigned long long n_reg;
  size_t nseg_out;
  int has_tail;
  unsigned long long tail_end;
};
cudaError_t launch_intgdump(const IdLaunch &p, cudaStream_t st);
const char *intgdump_path(const IdLaunch &p);

// Interpolating polyphase FIR on 16-bit samples (fused CIC interpolator + FIR cascade): upfir_q15.cu
struct UpLaunch {
  Fmt facc, fout;
  int R, taps_total, planes, lsh;
  uint32_t C;
  int interleaved;
  const void *in;        // n inputs per channel (int16 container)
  void *out;             // n_out outputs per channel, planar stride n_out
  size_t n, n_out;
  unsigned long long n_seen, out_first;
  const void *tail;      // [C][H] previous inputs
  int H;
  const uint32_t *cw;    // [C][upfir_q15_words] packed composite taps
};
bool upfir_q15_geometry(int R, int taps_total, int max_abs_bits);
int upfir_q15_planes(int max_abs_bits);
int upfir_q15_words(int R, int taps_total, int planes);
void upfir_q15_pack(const int64_t *c, int taps_total, int R, int planes, uint32_t *out);
cudaError_t launch_upfir_q15(const UpLaunch &p, cudaStream_t st);

// Weighted moving average over bursts (ac_mv_avg): mv_avg.cu
struct MvLaunch {
  Fmt fin, fcoeff, facc, fout;
  int taps, win;             // window span (odd), b2d_window_mode
  const void *in;            // whole bursts of n_sample samples
  void *out;                 // `per` outputs per burst
  const int64_t *coeff64;    // [taps]
  size_t n_sample, per, n_out;
};
cudaError_t launch_mvavg(const MvLaunch &p, cudaStream_t st);

// Packed host-link format (wire.cu): `count` values in 2 / 4 / 8-byte containers -> wire_bytes (< container) bytes each.
cudaError_t launch_pack_wire(const void *src, int container_bytes, void *dst, int wire_bytes, size_t count, cudaStream_t st);

}  // namespace b2d
