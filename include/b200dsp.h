/* b200dsp.h -- C-ABI of the B200-native fixed-point streaming-filter engine.
 *
 * This is the drop-in boundary for the FIR / CIC hot path of hlslibs/ac_dsp.  The reference has
 * no FFI layer: its boundary is five C++ class templates whose run() drains ac_channel FIFOs
 * (citations relative to the reference's include/ac_dsp/):
 *
 *   ac_fir_const_coeffs<IN,OUT,COEFF,ACC,N_TAPS,ftype>::run(in, out)             ac_fir_const_coeffs.h:309-355
 *   ac_fir_load_coeffs <IN,OUT,COEFF,ACC,N_TAPS,ftype>::run(in, coeffs, out, ld)  ac_fir_load_coeffs.h:300-365
 *   ac_fir_prog_coeffs <IN,OUT,COEFF,ACC,N_TAPS,ftype>::run(in, out, coeffs[])    ac_fir_prog_coeffs.h:261-303
 *   ac_cic_dec_full    <IN,OUT,R,M,N>::run(in, out)                              ac_cic_dec_full.h:147-222
 *   ac_cic_intr_full   <IN,OUT,R,M,N>::run(in, out)                              ac_cic_intr_full.h:137-215
 *
 * Here one opaque handle stands for one such object (times n_channels independent copies); the
 * template parameters become a run-time descriptor, ac_fixed<W,I,S,Q,O> values cross the boundary as
 * raw two's-complement integers, and run() takes plain arrays instead of channels.  The header
 * facade in include/b200dsp/ re-creates the five class templates on top of these entry points.
 *
 * Raw sample containers: the smallest of int16_t / int32_t / int64_t that holds W bits
 * (b2d_container_bytes), value sign- (S=1) or zero- (S=0) extended, little endian.
 *
 * Everything computes on the GPU (sm_100a).  There is no CPU fallback: a configuration the CUDA
 * kernels cannot reproduce bit-exactly is rejected with B2D_EUNSUPPORTED at create time.
 * No function throws; every function returns a b2d_status (0 = ok, negative = error).
 * A handle is single-caller (like the reference object); distinct handles are independent.
 */
#ifndef B200DSP_H
#define B200DSP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B2D_OK = 0,
  B2D_EUNSUPPORTED = -1, /* valid reference configuration the engine does not implement            */
  B2D_EINVAL = -2,       /* malformed argument (null pointer, zero taps, W out of range ...)       */
  B2D_ECUDA = -3,        /* CUDA runtime error (b2d_last_error() has the text)                     */
  B2D_ENCCL = -4,        /* NCCL error or NCCL library not loadable                                */
  B2D_ENOMEM = -5,
  B2D_ESTATE = -6        /* call sequence error: run() before coefficients were loaded, ...        */
} b2d_status;

/* ac_q_mode / ac_o_mode in the AC Datatypes order. */
typedef enum { B2D_TRN = 0, B2D_RND, B2D_TRN_ZERO, B2D_RND_ZERO, B2D_RND_INF, B2D_RND_MIN_INF, B2D_RND_CONV, B2D_RND_CONV_ODD } b2d_qmode;
typedef enum { B2D_WRAP = 0, B2D_SAT, B2D_SAT_ZERO, B2D_SAT_SYM } b2d_omode;

/* ac_fixed<W, I, S, Q, O>.  Supported: inputs / coefficients W <= 32; accumulators and outputs W <= 64 with every
 * quantisation and overflow mode.  Accumulators with Q in {TRN, RND} and O = WRAP make the per-tap `acc += a*b`
 * re-quantisation order-independent and run on the fast kernel families (b2d_*_path); saturating or sign-dependent
 * accumulators are evaluated tap by tap in the reference's own order by the generic kernels (slower, still on the GPU).
 * Combinations whose intermediates exceed 128 bits are rejected with B2D_EUNSUPPORTED at create time. */
typedef struct { int32_t W, I, S, Q, O; } b2d_fmt;

/* FTYPE enum of the reference (ac_fir_const_coeffs.h:96). The three hot classes dispatch the first
 * six; with an _ANTI value their run() leaves the output unwritten, so the engine rejects those. */
typedef enum {
  B2D_SHIFT_REG = 0, B2D_ROTATE_SHIFT, B2D_C_BUFF, B2D_FOLD_EVEN, B2D_FOLD_ODD, B2D_TRANSPOSED,
  B2D_FOLD_EVEN_ANTI, B2D_FOLD_ODD_ANTI
} b2d_ftype;

/* REG_SHARE = ac_fir_reg_share (ac_fir_reg_share.h:257-303): dispatches SHIFT_REG, FOLD_EVEN, FOLD_ODD and the two _ANTI
 * (anti-symmetric, pre-SUBTRACT) architectures, walks the taps upwards, reads its taps from a blocked coefficient RAM. */
typedef enum { B2D_FIR_CONST = 0, B2D_FIR_LOAD = 1, B2D_FIR_PROG = 2, B2D_FIR_REG_SHARE = 3 } b2d_fir_kind;
typedef enum { B2D_CIC_DEC = 0, B2D_CIC_INTR = 1 } b2d_cic_mode;

/* Multi-channel sample layout of the in/out arrays of one run() call with n samples per channel:
 * PLANAR: channel c at [c*n, (c+1)*n);  INTERLEAVED: sample i of channel c at [i*n_channels + c]
 * (16-bit IQ = 2 interleaved channels sharing one coefficient set). */
typedef enum { B2D_PLANAR = 0, B2D_INTERLEAVED = 1 } b2d_layout;

/* Format of the OUTPUT array of the host-buffer run() calls (b2d_*_run; the _run_dev calls always use containers).
 * CONTAINER (default): one int16 / int32 / int64 container per value, as everywhere else.
 * PACKED: ceil(W_out / 8) little-endian bytes per value, same element order -- an ac_fixed<40,8> result is 5 bytes
 * instead of 8.  run() on host memory is bound by the host link, so fewer bytes per value is more samples per second;
 * b2d_unpack_wire() widens a packed array to containers on the host.  b2d_wire_bytes() gives the bytes per value. */
typedef enum { B2D_WIRE_CONTAINER = 0, B2D_WIRE_PACKED = 1 } b2d_wire;

typedef struct {
  b2d_fmt in, coeff, acc, out;  /* IN_TYPE, COEFF_TYPE, ACC_TYPE, OUT_TYPE                           */
  uint32_t n_taps;              /* N_TAPS >= 1                                                      */
  int32_t ftype;                /* b2d_ftype                                                        */
  int32_t kind;                 /* b2d_fir_kind: which reference class's load protocol is mirrored  */
  uint32_t n_channels;          /* independent filter instances sharing this descriptor (>= 1)      */
  int32_t layout;               /* b2d_layout                                                       */
  int32_t device;               /* CUDA device ordinal, -1 = current device                         */
} b2d_fir_desc;

typedef struct {
  b2d_fmt in, out;              /* IN_TYPE, OUT_TYPE; the lossless INT_TYPE is derived internally   */
  uint32_t R, M, N;             /* rate change 2..256 (R = 1: B2D_EUNSUPPORTED), differential delay >= 1, stages 1..16 */
  int32_t mode;                 /* b2d_cic_mode                                                     */
  uint32_t n_channels;
  int32_t layout;
  int32_t device;
} b2d_cic_desc;

typedef struct b2d_fir b2d_fir;
typedef struct b2d_cic b2d_cic;
typedef struct b2d_comm b2d_comm;

/* ---- library ---------------------------------------------------------------------------- */
const char *b2d_version(void);
const char *b2d_strerror(int status);
const char *b2d_last_error(void);                 /* thread-local detail text of the last failure   */
int b2d_container_bytes(int32_t W);               /* 2, 4 or 8                                      */
int b2d_device_count(void);                       /* number of visible CUDA devices (0 if none)     */
/* Page-locked host buffers: run() on HOST memory overlaps its copies with compute only when the
 * buffers are page-locked (these, or any cudaHostAlloc / cudaHostRegister / torch pinned memory). */
int b2d_host_alloc(void **p, size_t bytes);
int b2d_host_free(void *p);
int b2d_wire_bytes(int32_t W, int32_t wire);      /* bytes per value of a W-bit format in a host output array */
/* `count` packed values (B2D_WIRE_PACKED) of an ac_fixed<W,.,S> format -> containers (b2d_container_bytes(W) each), host side. */
int b2d_unpack_wire(const void *packed, size_t count, int32_t W, int32_t S, void *out_containers);

/* ---- FIR: ac_fir_const_coeffs / ac_fir_load_coeffs / ac_fir_prog_coeffs -------------------- */
/* Class instantiation + constructor: zeroed delay line (ac_fir_load_coeffs.h:134-139). */
int b2d_fir_create(b2d_fir **h, const b2d_fir_desc *desc);
int b2d_fir_destroy(b2d_fir *h);
/* Coefficients, n == n_taps raw values in the coefficient container, HOST memory.
 *   CONST: the constructor's pointer (ac_fir_const_coeffs.h:314) -- allowed once.
 *   LOAD : the ld=true phase of run() (ac_fir_load_coeffs.h:324-331).
 *   PROG : the array argument of run() (ac_fir_prog_coeffs.h:277); may change between run() calls,
 *          the delay line is kept.  B2D_TRANSPOSED keeps ACC_TYPE partial sums, not samples
 *          (ac_fir_load_coeffs.h:265-278, ac_fir_prog_coeffs.h:232-247): after a change the reference's next
 *          n_taps-1 outputs are old-tap partial sums plus new-tap products, and so are the engine's (the history
 *          is converted into those pending sums at the change).
 * channel = -1 loads every channel.  With a communicator attached (b2d_fir_set_comm) the values of
 * rank `root` are broadcast to all ranks with one ncclBroadcast; other ranks may pass NULL. */
int b2d_fir_load(b2d_fir *h, const void *coeff_raw, size_t n, int32_t channel);
int b2d_fir_set_comm(b2d_fir *h, b2d_comm *comm, int32_t root);
/* ac_fir_reg_share's coefficient addressing (ac_fir_reg_share.h:122-133): tap t of the architecture's tap loop reads
 * ram[(t / blk_sz) * mem_word_width + blk_offset + t % blk_sz]; n_ram words of raw COEFF_TYPE values in HOST memory.
 * The tap loop length (N_TAPS, N_TAPS/2 or (N_TAPS-1)/2+1 by ftype) must be a multiple of blk_sz -- otherwise the
 * reference indexes its delay line out of range.  With mem_word_width = blk_sz = 1, blk_offset = 0 this is b2d_fir_load. */
int b2d_fir_load_blocked(b2d_fir *h, const void *coeff_ram, size_t n_ram, uint32_t mem_word_width, uint32_t blk_sz,
                         uint32_t blk_offset, int32_t channel);
/* ac_firProgCoeffs_delay_line (ac_fir_reg_share.h:304-307): OUT_TYPE(reg[N_TAPS-1]) -- the sample leaving the delay line,
 * one raw value per channel in the output container (REG_SHARE handles only). */
int b2d_fir_delay_line_out(b2d_fir *h, void *out_raw);
/* One output per channel from an explicit delay line: window[c * n_taps + i] = reg[i] of channel c (reg[0] newest), raw
 * IN_TYPE values in HOST memory; out_raw[c] in the output container.  Does not read or change the handle's own delay
 * line.  This is ac_fir_reg_share's calling convention -- the delay line belongs to the caller and run() is scalar
 * (ac_fir_reg_share.h:262-303) -- and what its facade class uses; block processing goes through b2d_fir_run. */
int b2d_fir_run_window(b2d_fir *h, const void *window, void *out_raw);
/* run() sample loop: n samples per channel in, n per channel out (*n_out = n).
 * _run takes HOST buffers (copies in, computes, copies out, synchronous); _run_dev takes DEVICE
 * buffers and is asynchronous on `cuda_stream` (a cudaStream_t, NULL = default stream). */
int b2d_fir_run(b2d_fir *h, const void *in, size_t n, void *out, size_t *n_out);
int b2d_fir_run_dev(b2d_fir *h, const void *d_in, size_t n, void *d_out, size_t *n_out, void *cuda_stream);
int b2d_fir_set_wire(b2d_fir *h, int32_t wire);   /* b2d_wire: format of b2d_fir_run's output array          */
int b2d_fir_reset(b2d_fir *h);                    /* back to the constructed state, coefficients kept */
/* Checkpoint: the delay line etc. as an opaque blob (size via _state_bytes). */
int b2d_fir_state_bytes(b2d_fir *h, size_t *bytes);
int b2d_fir_get_state(b2d_fir *h, void *blob, size_t bytes);
int b2d_fir_set_state(b2d_fir *h, const void *blob, size_t bytes);
/* Name of the kernel family chosen for this descriptor: "fir_q15" (16-bit operands, DP2A byte planes), "fir_q24"
 * (samples of 17..24 bits, taps <= 16 bits: DP2A on three sample byte planes), "fir_wide" (operands <= 32 bits, wrapping
 * 64-bit accumulator) or "fir_generic" (every format and mode, reference tap order).  A "fir_q15" filter of 96..2049 taps
 * reports "fir_ovs" once the loaded coefficients allow the overlap-save evaluation (blocks of 4096 samples through an FP64
 * FFT whose a-priori error bound for THESE taps is below 1/2, so the rounded result is the exact integer sum); calls
 * too short to repay a round of blocks (under about 5 * 10^5 IQ samples at 256 taps, 2.5 * 10^4 per channel at 1024 taps on
 * eight channels) still run "fir_q15".  B2D_FIR_OVS=0 in the
 * environment at create time turns it off. */
const char *b2d_fir_path(b2d_fir *h);
/* Overlap-save diagnostics: `bound` = the a-priori bound of |FP64 result - exact sum| for the loaded taps (armed below
 * 0.49); `resid` = the largest distance of a result from an integer seen so far when B2D_OVS_RESID=1 was set at create
 * time, else -1.  Either pointer may be NULL. */
int b2d_fir_ovs_margin(b2d_fir *h, double *bound, double *resid);

/* ---- CIC: ac_cic_dec_full / ac_cic_intr_full ---------------------------------------------- */
/* Note on M > 2: the reference's comb shifts its delay line with an ascending copy loop (ac_cic_full_core.h:247-251),
 * which makes the differential delay min(M, 2) while the lossless internal width still grows with M; the engine
 * reproduces exactly that. */
int b2d_cic_create(b2d_cic **h, const b2d_cic_desc *desc);
int b2d_cic_destroy(b2d_cic *h);
/* Lossless internal width of find_inter_type_cic_dec / _intr (ac_cic_dec_full.h:132, ac_cic_intr_full.h:122). */
int b2d_cic_int_width(const b2d_cic_desc *desc, int32_t *outW);
/* Upper bound on outputs per channel produced by a run() of n inputs per channel. */
size_t b2d_cic_max_out(b2d_cic *h, size_t n);
/* run(): DEC emits the samples at global input indices 0, R, 2R, ... ; INTR emits, after K inputs in
 * total, max(0, (K-1)R + 1 - (N-1)) outputs in total (the last input of a call yields one output until
 * more data arrives -- the reference's stream-edge rule).  *n_out = outputs per channel of this call;
 * PLANAR outputs use a channel stride of *n_out. */
int b2d_cic_run(b2d_cic *h, const void *in, size_t n, void *out, size_t *n_out);
int b2d_cic_run_dev(b2d_cic *h, const void *d_in, size_t n, void *d_out, size_t *n_out, void *cuda_stream);
int b2d_cic_set_wire(b2d_cic *h, int32_t wire);
int b2d_cic_reset(b2d_cic *h);
int b2d_cic_state_bytes(b2d_cic *h, size_t *bytes);
int b2d_cic_get_state(b2d_cic *h, void *blob, size_t bytes);
int b2d_cic_set_state(b2d_cic *h, const void *blob, size_t bytes);
const char *b2d_cic_path(b2d_cic *h);

/* ---- cascade: ac_cic_intr_full -> ac_fir_* as ONE object ------------------------------------ */
/* BASELINE config 5 (CIC interpolator followed by a FIR).  In the reference these are two objects joined by an
 * ac_channel:  cic.run(in, mid); fir.run(mid, out);  (ac_cic_intr_full.h:150-153, ac_fir_const_coeffs.h:321-355).
 * A b2d_cicfir handle stands for that pair and produces bit-identical results; `cic` must be an INTR descriptor,
 * fir->in must equal cic->out, both must have the same n_channels.  Input layout = cic->layout, outputs are PLANAR
 * with a channel stride of *n_out (like b2d_cic_run).  When neither stage drops bits (the lossless CIC type is
 * passed on unchanged and the FIR accumulator is an exact-shift AC_WRAP one) both stages run as a single polyphase
 * kernel on the original 16-bit samples and the intermediate stream never reaches HBM (b2d_cicfir_path:
 * "cicfir_fused"); otherwise the two kernels run back to back through a device buffer ("cicfir_two_stage"). */
typedef struct b2d_cicfir b2d_cicfir;
int b2d_cicfir_create(b2d_cicfir **h, const b2d_cic_desc *cic, const b2d_fir_desc *fir);
int b2d_cicfir_destroy(b2d_cicfir *h);
int b2d_cicfir_load(b2d_cicfir *h, const void *coeff_raw, size_t n, int32_t channel);   /* the FIR's taps, as b2d_fir_load */
size_t b2d_cicfir_max_out(b2d_cicfir *h, size_t n);
int b2d_cicfir_run(b2d_cicfir *h, const void *in, size_t n, void *out, size_t *n_out);
int b2d_cicfir_run_dev(b2d_cicfir *h, const void *d_in, size_t n, void *d_out, size_t *n_out, void *cuda_stream);
int b2d_cicfir_set_wire(b2d_cicfir *h, int32_t wire);
int b2d_cicfir_reset(b2d_cicfir *h);
const char *b2d_cicfir_path(b2d_cicfir *h);
int b2d_cicfir_state_bytes(b2d_cicfir *h, size_t *bytes);   /* checkpoint, as b2d_fir_get_state */
int b2d_cicfir_get_state(b2d_cicfir *h, void *blob, size_t bytes);
int b2d_cicfir_set_state(b2d_cicfir *h, const void *blob, size_t bytes);

/* ---- polyphase decimator: ac_poly_dec ----------------------------------------------------------- */
/* ac_poly_dec<IN, COEFF, STR_COEFF, ACC, OUT, NTAPS, DF>::run(data_in, data_out, coeffs_st)  (ac_poly_dec.h:87-137):
 * one output per DF inputs, out[m] = sum_{r<DF} sum_{tp<NTAPS} coeffs[tp + NTAPS*r] * x[(m - tp)*DF + DF-1 - r], every
 * `+=` re-quantised to ACC_TYPE.  Coefficients: n_taps*df raw values in the reference's phase order (the coeffs[] member
 * of its coefficient struct).  A call consumes whole groups of DF samples; up to DF-1 trailing samples stay pending
 * inside the handle (in the reference they stay queued on data_in).  Outputs are PLANAR with a channel stride of *n_out. */
typedef struct {
  b2d_fmt in, coeff, acc, out;  /* IN_TYPE, COEFF_TYPE, ACC_TYPE, OUT_TYPE                            */
  uint32_t n_taps;              /* NTAPS: taps per phase (NTAPS * DF coefficients in all)             */
  uint32_t df;                  /* DF: decimation factor >= 1                                         */
  uint32_t n_channels;
  int32_t layout;               /* layout of the input                                                */
  int32_t device;
} b2d_polydec_desc;
typedef struct b2d_polydec b2d_polydec;
int b2d_polydec_create(b2d_polydec **h, const b2d_polydec_desc *desc);
int b2d_polydec_destroy(b2d_polydec *h);
int b2d_polydec_load(b2d_polydec *h, const void *coeff_raw, size_t n, int32_t channel);
size_t b2d_polydec_max_out(b2d_polydec *h, size_t n);
int b2d_polydec_run(b2d_polydec *h, const void *in, size_t n, void *out, size_t *n_out);
int b2d_polydec_run_dev(b2d_polydec *h, const void *d_in, size_t n, void *d_out, size_t *n_out, void *cuda_stream);
int b2d_polydec_set_wire(b2d_polydec *h, int32_t wire);
int b2d_polydec_reset(b2d_polydec *h);
const char *b2d_polydec_path(b2d_polydec *h);
int b2d_polydec_state_bytes(b2d_polydec *h, size_t *bytes);
int b2d_polydec_get_state(b2d_polydec *h, void *blob, size_t bytes);
int b2d_polydec_set_state(b2d_polydec *h, const void *blob, size_t bytes);

/* ---- polyphase interpolating FIR: ac_poly_intr ---------------------------------------------------- */
/* ac_poly_intr<IN, COEFF, ACC, OUT, STR_CTRL, STR_COEFF, NTAPS, COEFFSZ, IF, ftype>::run(data_in, data_out, ctrl_st,
 * coeffs_st, read_ctrl_chan)  (ac_poly_intr.h:261-312).  A reference run() call consumes one read_ctrl token: true loads
 * the control struct {bool sign[IF]; ac_int<8,false> corr[IF];} and the coefficient struct (b2d_polyintr_load), false
 * consumes one sample and writes IF outputs (b2d_polyintr_run with n samples = n such calls).  NTAPS is the length of
 * the low-rate delay line.  ftype is the enum of ac_poly_intr.h:95: the two folded architectures implement the
 * symmetric-pair technique and emit the outputs of a step one step late (nothing for the very first sample, :160);
 * FOLD_ANTI is the plain polyphase form and ignores sign / corr.  Coefficients per channel: IF * NTAPS/2 (FOLD_EVEN,
 * index i + j*NTAPS/2), IF * (NTAPS/2 + 1) (FOLD_ODD, i + (NTAPS/2+1)*j), IF * NTAPS (FOLD_ANTI, i + NTAPS*j).
 * Outputs: IF per step, phase-minor, planar over channels with stride *n_out. */
typedef enum { B2D_PI_FOLD_EVEN = 0, B2D_PI_FOLD_ODD = 1, B2D_PI_FOLD_ANTI = 2 } b2d_polyintr_ftype;
typedef struct {
  b2d_fmt in, coeff, acc, out;  /* IN_TYPE, COEFF_TYPE, ACC_TYPE, OUT_TYPE                              */
  uint32_t n_taps;              /* NTAPS                                                                */
  uint32_t intr_factor;         /* IF, 1 .. 255 (corr is ac_int<8,false> in the reference's usage)      */
  int32_t ftype;                /* b2d_polyintr_ftype                                                   */
  uint32_t n_channels;
  int32_t layout;               /* layout of the input                                                  */
  int32_t device;
} b2d_polyintr_desc;
typedef struct b2d_polyintr b2d_polyintr;
int b2d_polyintr_create(b2d_polyintr **h, const b2d_polyintr_desc *desc);
int b2d_polyintr_destroy(b2d_polyintr *h);
size_t b2d_polyintr_coeffsz(b2d_polyintr *h);
/* coeff_raw: b2d_polyintr_coeffsz(h) raw coefficients; sign / corr: IF bytes each (NULL = all true / identity);
 * channel -1 = all channels.  May be called between runs: the accumulators of the last step keep the values they were
 * computed with, the new sign / corr apply to the outputs that follow (as in the reference). */
int b2d_polyintr_load(b2d_polyintr *h, const void *coeff_raw, size_t n, const uint8_t *sign, const uint8_t *corr, int32_t channel);
size_t b2d_polyintr_max_out(b2d_polyintr *h, size_t n);
int b2d_polyintr_run(b2d_polyintr *h, const void *in, size_t n, void *out, size_t *n_out);
int b2d_polyintr_run_dev(b2d_polyintr *h, const void *d_in, size_t n, void *d_out, size_t *n_out, void *cuda_stream);
int b2d_polyintr_set_wire(b2d_polyintr *h, int32_t wire);
int b2d_polyintr_reset(b2d_polyintr *h);
/* "polyintr_q15" (DP2A polyphase kernel, FOLD_ANTI on 16-bit operands), "polyintr_wide" (64-bit modular), "polyintr_generic" */
const char *b2d_polyintr_path(b2d_polyintr *h);
/* checkpoint: delay line, the parked accumulators acc_a / acc_b of the last step and `init` (ac_poly_intr.h:107-111) */
int b2d_polyintr_state_bytes(b2d_polyintr *h, size_t *bytes);
int b2d_polyintr_get_state(b2d_polyintr *h, void *blob, size_t bytes);
int b2d_polyintr_set_state(b2d_polyintr *h, const void *blob, size_t bytes);

/* ---- integrate and dump: ac_intg_dump ------------------------------------------------------------ */
/* ac_intg_dump<IN, ACC, OUT, N_TYPE, NS, CHN>::run(data_in, data_out, n_sample)  (ac_intg_dump.h:113-151).  One call =
 * n_frames frames; frame f reads the token n_sample[f]: with 1 <= n_sample[f] <= NS it consumes n_sample[f] samples of each
 * of the CHN interleaved channels and dumps CHN sums, otherwise it consumes NS samples per channel, dumps nothing and
 * the running sums carry on into the next frame (and the next call).  n_in must be exactly what the frames consume --
 * the reference would read past the end of its input channel otherwise.  out: CHN values per dumping frame, frame-major. */
typedef struct {
  b2d_fmt in, acc, out;   /* IN_TYPE, ACC_TYPE, OUT_TYPE                                                     */
  uint32_t ns, chn;       /* NS: longest frame; CHN: interleaved channels                                    */
  int32_t device;
} b2d_intgdump_desc;
typedef struct b2d_intgdump b2d_intgdump;
int b2d_intgdump_create(b2d_intgdump **h, const b2d_intgdump_desc *desc);
int b2d_intgdump_destroy(b2d_intgdump *h);
int b2d_intgdump_run(b2d_intgdump *h, const void *in, size_t n_in, const uint32_t *n_sample, size_t n_frames, void *out, size_t *n_out);
int b2d_intgdump_run_dev(b2d_intgdump *h, const void *d_in, size_t n_in, const uint32_t *n_sample, size_t n_frames, void *d_out,
                         size_t *n_out, void *cuda_stream);
int b2d_intgdump_reset(b2d_intgdump *h);
/* kernel family the last run took: "intgdump_vec" (128-bit loads, equal frames), "intgdump_warp", "intgdump_thread" */
const char *b2d_intgdump_path(b2d_intgdump *h);
/* checkpoint: the running sums temp[CHN] (ac_intg_dump.h:78) */
int b2d_intgdump_state_bytes(b2d_intgdump *h, size_t *bytes);
int b2d_intgdump_get_state(b2d_intgdump *h, void *blob, size_t bytes);
int b2d_intgdump_set_state(b2d_intgdump *h, const void *blob, size_t bytes);

/* ---- weighted moving average: ac_mv_avg ------------------------------------------------------------- */
/* ac_mv_avg<MAX_SAMPLE, TAPS, WIN_TYPE, IN, OUT, ACC, COEFF, S_TYPE>::run(data_in, data_out, n_sample)  (ac_mv_avg.h:140-204):
 * the input is processed in bursts of n_sample samples (TAPS <= n_sample <= MAX_SAMPLE); within a burst
 * out[i] = sum_{j=-TAPS/2..TAPS/2} (ACC_TYPE) x~[i+j] * coeffs[j + TAPS/2], the sum re-quantised to ACC_TYPE at every tap.
 * x~ extends the burst at its two ends: B2D_CLIP repeats the edge sample, B2D_MIRROR reflects about it (n_sample outputs
 * per burst); B2D_WIN has no boundary processing and emits only the n_sample - TAPS + 1 points whose window is full.
 * The weights are the constructor's const array (TAPS raw COEFF_TYPE values, HOST memory).  No state survives a run().
 * PARITY UNPINNED: the window semantics come from ac_math's ac_window.h, which neither the reference tree nor this
 * image contains and for which the reference ships no vector; they are restated from the manual. */
typedef enum { B2D_WIN = 0, B2D_CLIP = 1, B2D_MIRROR = 2 } b2d_window_mode;
typedef struct {
  b2d_fmt in, out, acc, coeff;   /* IN_TYPE, OUT_TYPE, ACC_TYPE, COEFF_TYPE                                       */
  uint32_t max_sample, taps;     /* MAX_SAMPLE; TAPS (odd)                                                        */
  int32_t win_type;              /* b2d_window_mode                                                               */
  int32_t device;
} b2d_mvavg_desc;
typedef struct b2d_mvavg b2d_mvavg;
int b2d_mvavg_create(b2d_mvavg **h, const b2d_mvavg_desc *desc, const void *coeff_raw);
int b2d_mvavg_destroy(b2d_mvavg *h);
size_t b2d_mvavg_max_out(b2d_mvavg *h, size_t n);
/* n_in samples = whole bursts of n_sample; *n_out = outputs of all bursts, burst-major. */
int b2d_mvavg_run(b2d_mvavg *h, const void *in, size_t n_in, size_t n_sample, void *out, size_t *n_out);
int b2d_mvavg_run_dev(b2d_mvavg *h, const void *d_in, size_t n_in, size_t n_sample, void *d_out, size_t *n_out, void *cuda_stream);
const char *b2d_mvavg_path(b2d_mvavg *h);

/* ---- multi-GPU: one process per GPU, channels sharded, coefficients broadcast once --------- */
#define B2D_UNIQUE_ID_BYTES 128
/* Channel c of n_channels lives on rank c % world (contiguous block alternative: see DESIGN.md). */
int b2d_shard_count(uint32_t n_channels, int32_t rank, int32_t world, uint32_t *n_local);
int b2d_comm_unique_id(void *id128);                                 /* rank 0: ncclGetUniqueId     */
int b2d_comm_create(b2d_comm **c, const void *id128, int32_t rank, int32_t world, int32_t device);
int b2d_comm_destroy(b2d_comm *c);
int b2d_comm_barrier(b2d_comm *c);                                   /* 1-element all-reduce + sync */

#ifdef __cplusplus
}
#endif
#endif /* B200DSP_H */
