"""oracle/ref_configs.py -- TEST INFRASTRUCTURE.

The reference classes are C++ templates, so every (formats, taps, R/M/N) combination the
parity tests want from the *real* reference has to be instantiated at compile time.  This
table is the single source of truth for those instantiations: `gen_ref_cfgs.py` turns it
into the X-macro include files compiled into oracle/_ref/libacdsp_ref.so, and
oracle/refdrv.py uses it to look a configuration up by value.

fmt tuples are ac_fixed<W, I, S, Q, O> with Q/O given by name.
"""
import math

TRN, RND = "AC_TRN", "AC_RND"
WRAP = "AC_WRAP"


def fmt(W, I, S=True, Q=TRN, O=WRAP):
    return (W, I, bool(S), Q, O)


# (name, in, coeff, acc, out, taps-list)
FIR_FORMATS = [
    # BASELINE.json configs 1/2/4: <16,1> x <16,1> -> <40,8>
    ("q15_acc40", fmt(16, 1), fmt(16, 1), fmt(40, 8), fmt(40, 8), [1, 2, 3, 16, 27, 64, 256, 1024]),
    # reference bench formats (tests/rtest_ac_fir_{const,load,prog}_coeffs.cpp)
    ("bench_const", fmt(16, 8), fmt(32, 16), fmt(64, 32), fmt(64, 32), [29]),
    ("bench_load", fmt(32, 16), fmt(32, 16), fmt(64, 32), fmt(64, 32), [27, 8]),
    ("bench_prog", fmt(28, 6), fmt(23, 7), fmt(64, 32), fmt(64, 32), [27, 12]),
    # per-tap truncation (F_acc < F_in + F_c), narrow output: exercises floor + wrap everywhere
    ("q15_trunc", fmt(16, 1), fmt(16, 1), fmt(24, 4), fmt(16, 1), [5, 16, 63]),
    # rounding accumulator / output
    ("q15_rnd", fmt(16, 1), fmt(16, 1), fmt(24, 4, True, RND), fmt(16, 1, True, RND), [16, 31]),
    # unsigned input, signed coefficients
    ("u12", fmt(12, 0, False), fmt(14, 2), fmt(30, 6), fmt(20, 4), [9, 32]),
    # BASELINE.json config 5 second stage: <20,5> x <16,1> -> <40,8>, 63 taps
    ("cic_post", fmt(20, 5), fmt(16, 1), fmt(40, 8), fmt(40, 8), [63]),
    # accumulator too narrow for the fold pre-add (wrap inside FOLD_ODD's `fold`)
    ("narrow_acc", fmt(16, 1), fmt(16, 1), fmt(18, 1), fmt(18, 1), [7, 10]),
    # order-dependent accumulators: saturation / sign-dependent rounding (the reference's tap order matters)
    ("q15_sat", fmt(16, 1), fmt(16, 1), fmt(24, 4, True, TRN, "AC_SAT"), fmt(16, 1, True, "AC_RND_CONV", "AC_SAT_SYM"), [9, 16]),
    ("q15_tz", fmt(16, 1), fmt(16, 1), fmt(30, 6, True, "AC_TRN_ZERO", "AC_SAT_ZERO"), fmt(12, 1, True, "AC_RND_INF", "AC_SAT"), [10]),
    # unsigned ACC_TYPE: FOLD_ODD's `fold = ACC_TYPE(a + b)` wraps every negative pre-add before the multiply
    ("uacc", fmt(16, 1), fmt(16, 1), fmt(40, 8, False), fmt(40, 8, False), [9, 16]),
]


def log2_ceil(n):
    return 0 if n <= 1 else (n - 1).bit_length()


def cic_int_width(mode, W, S, R, M, N):
    """Lossless internal width: ac_cic_dec_full.h:132 / ac_cic_intr_full.h:122."""
    g = (R ** N) * (M ** N) if mode == "dec" else (R ** (N - 1)) * (M ** N)
    return log2_ceil(g) + W + (0 if S else 1)


def _cic_list():
    out = []
    for mode in ("dec", "intr"):
        # sweep with <16,1> input and the lossless output type
        for R in (2, 3, 4, 7, 8, 16):
            for M in (1, 2):
                for N in (1, 2, 3, 4, 5):
                    W, I = 16, 1
                    ow = cic_int_width(mode, W, True, R, M, N)
                    if ow > 64:
                        continue
                    out.append((mode, R, M, N, fmt(W, I), fmt(ow, ow - (W - I))))
        # narrowed outputs (INT_TYPE -> OUT_TYPE conversion with truncation and wrap)
        out.append((mode, 8, 1, 4, fmt(16, 1), fmt(16, 1)))
        out.append((mode, 4, 2, 3, fmt(16, 1), fmt(12, 4, True, RND)))
        # unsigned input
        out.append((mode, 5, 1, 3, fmt(10, 2, False), fmt(24, 12)))
    # the two golden-vector configurations (tests/ac_cic_{dec,intr}_full_param.h:34-47)
    out.append(("dec", 7, 2, 4, fmt(32, 16), fmt(48, 32)))
    out.append(("intr", 7, 2, 5, fmt(32, 16), fmt(49, 33)))
    # wide state (> 32 bits) with 16-bit input
    out.append(("dec", 16, 2, 5, fmt(16, 1), fmt(41, 26)))
    out.append(("dec", 32, 1, 6, fmt(16, 1), fmt(46, 31)))
    out.append(("intr", 32, 1, 6, fmt(16, 1), fmt(41, 26)))
    seen, uniq = set(), []
    for c in out:
        if c not in seen:
            seen.add(c)
            uniq.append(c)
    return uniq


CIC_CONFIGS = _cic_list()


# ac_fir_reg_share instantiations:
# (N_TAPS, in, out, coeff, acc, MEM_WORD_WIDTH, BLK_SZ, BLK_OFFSET, ftype)
_Q15, _ACC40 = fmt(16, 1), fmt(40, 8)
RS_CONFIGS = [
    (16, _Q15, _ACC40, _Q15, _ACC40, 1, 1, 0, "SHIFT_REG"),
    (16, _Q15, _ACC40, _Q15, _ACC40, 1, 1, 0, "FOLD_EVEN"),
    (16, _Q15, _ACC40, _Q15, _ACC40, 1, 1, 0, "FOLD_EVEN_ANTI"),
    (15, _Q15, _ACC40, _Q15, _ACC40, 1, 1, 0, "FOLD_ODD"),
    (15, _Q15, _ACC40, _Q15, _ACC40, 1, 1, 0, "FOLD_ODD_ANTI"),
    (64, _Q15, _ACC40, _Q15, _ACC40, 1, 1, 0, "FOLD_EVEN_ANTI"),
    # blocked coefficient RAM: tap t reads ram[(t / BLK_SZ) * MEM_WORD_WIDTH + BLK_OFFSET + t % BLK_SZ]
    (16, _Q15, _ACC40, _Q15, _ACC40, 4, 2, 1, "SHIFT_REG"),
    (24, _Q15, _ACC40, _Q15, _ACC40, 8, 4, 4, "FOLD_EVEN_ANTI"),
    # per-tap truncation + narrow output (the anti-symmetric pre-add is inside the truncated product)
    (16, _Q15, fmt(16, 1), _Q15, fmt(24, 4), 1, 1, 0, "FOLD_EVEN_ANTI"),
    (15, _Q15, fmt(16, 1), _Q15, fmt(24, 4), 1, 1, 0, "FOLD_ODD_ANTI"),
    (16, _Q15, fmt(16, 1, True, RND), _Q15, fmt(24, 4, True, RND), 1, 1, 0, "FOLD_EVEN"),
    # the reference prog bench's formats, anti-symmetric
    (27, fmt(28, 6), fmt(64, 32), fmt(23, 7), fmt(64, 32), 1, 1, 0, "FOLD_ODD_ANTI"),
    (12, fmt(32, 16), fmt(64, 32), fmt(32, 16), fmt(64, 32), 1, 1, 0, "FOLD_EVEN_ANTI"),
    # unsigned samples: the difference of two unsigned values is signed
    (9, fmt(12, 0, False), fmt(20, 4), fmt(14, 2), fmt(30, 6), 1, 1, 0, "FOLD_ODD_ANTI"),
    # order-dependent accumulator: ac_fir_reg_share walks the taps upwards (ac_fir_reg_share.h:122-133)
    (16, _Q15, fmt(16, 1, True, "AC_RND_CONV", "AC_SAT_SYM"), _Q15, fmt(24, 4, True, TRN, "AC_SAT"), 1, 1, 0, "SHIFT_REG"),
    (10, _Q15, fmt(12, 1, True, "AC_RND_INF", "AC_SAT"), _Q15, fmt(30, 6, True, "AC_TRN_ZERO", "AC_SAT_ZERO"), 1, 1, 0, "FOLD_EVEN_ANTI"),
    # accumulator too narrow for the pre-add (wrap inside FOLD_ODD_ANTI's `fold`)
    (7, _Q15, fmt(18, 1), _Q15, fmt(18, 1), 1, 1, 0, "FOLD_ODD_ANTI"),
    # unsigned COEFF_TYPE with an _ANTI fold: the mirrored taps are negative, the effective taps signed
    (16, _Q15, _ACC40, fmt(14, 1, False), _ACC40, 1, 1, 0, "FOLD_EVEN_ANTI"),
    (15, _Q15, _ACC40, fmt(12, 0, False), _ACC40, 1, 1, 0, "FOLD_ODD_ANTI"),
    # everything unsigned: the pre-subtract wraps in the unsigned ACC_TYPE `fold`
    (9, fmt(12, 0, False), fmt(30, 6, False), fmt(14, 2, False), fmt(30, 6, False), 1, 1, 0, "FOLD_ODD_ANTI"),
]


# ac_poly_dec instantiations: (in, coeff, acc, out, NTAPS, DF)
PD_CONFIGS = [
    (_Q15, _Q15, _ACC40, _ACC40, 8, 2),
    (_Q15, _Q15, _ACC40, _ACC40, 16, 4),
    (_Q15, _Q15, _ACC40, _ACC40, 32, 8),        # 256 taps in all: the DDC partner of the R = 8 CIC decimator
    (_Q15, _Q15, _ACC40, _ACC40, 5, 3),
    (_Q15, _Q15, _ACC40, _ACC40, 1, 2),
    (fmt(28, 13), _Q15, fmt(48, 16), fmt(48, 16), 12, 2),          # the CIC decimator's <28,13> output as input
    (_Q15, _Q15, fmt(24, 4), fmt(16, 1), 16, 4),                  # per-tap truncation, narrow output
    (_Q15, _Q15, fmt(24, 4, True, RND), fmt(16, 1, True, RND), 6, 5),
    (fmt(12, 0, False), fmt(14, 2), fmt(30, 6), fmt(20, 4), 7, 3),  # unsigned samples
    (fmt(32, 16), fmt(32, 16), fmt(64, 32), fmt(64, 32), 9, 4),
    # order-dependent accumulators: taps upwards inside a phase, phases DF-1 .. 0 (ac_poly_dec.h:113-126)
    (_Q15, _Q15, fmt(24, 4, True, TRN, "AC_SAT"), fmt(16, 1, True, "AC_RND_CONV", "AC_SAT_SYM"), 8, 4),
    (_Q15, _Q15, fmt(30, 6, True, "AC_TRN_ZERO", "AC_SAT_ZERO"), fmt(12, 1, True, "AC_RND_INF", "AC_SAT"), 5, 2),
]


# ac_poly_intr instantiations: (in, coeff, acc, out, NTAPS, IF, ftype)
# ftype is the polyphase enum of ac_poly_intr.h:95 (FOLD_EVEN / FOLD_ODD: symmetric-pair structures; FOLD_ANTI: the plain
# polyphase form).  NTAPS = taps per phase (length of the low-rate delay line).
PI_FTYPES = ["FOLD_EVEN", "FOLD_ODD", "FOLD_ANTI"]
PI_CONFIGS = [
    (_Q15, _Q15, _ACC40, _ACC40, 8, 4, "FOLD_EVEN"),
    (_Q15, _Q15, _ACC40, _ACC40, 7, 4, "FOLD_ODD"),
    (_Q15, _Q15, _ACC40, _ACC40, 16, 4, "FOLD_ANTI"),            # 64 taps in all: the DUC partner of the R = 4 CIC interpolator
    (_Q15, _Q15, _ACC40, _ACC40, 32, 8, "FOLD_ANTI"),
    (_Q15, _Q15, _ACC40, _ACC40, 2, 3, "FOLD_EVEN"),
    (_Q15, _Q15, _ACC40, _ACC40, 1, 2, "FOLD_ODD"),
    (_Q15, _Q15, _ACC40, _ACC40, 5, 1, "FOLD_ANTI"),
    (fmt(32, 16), fmt(32, 16), fmt(64, 32), fmt(64, 32), 6, 3, "FOLD_EVEN"),   # the header's usage example (ac_poly_intr.h:46-49)
    (fmt(32, 16), fmt(32, 16), fmt(64, 32), fmt(64, 32), 5, 2, "FOLD_ODD"),
    (_Q15, _Q15, fmt(24, 4), fmt(16, 1), 8, 4, "FOLD_EVEN"),                  # fold and product both truncate
    (_Q15, _Q15, fmt(24, 4, True, RND), fmt(16, 1, True, RND), 9, 3, "FOLD_ODD"),
    (_Q15, _Q15, fmt(24, 4), fmt(16, 1), 6, 5, "FOLD_ANTI"),
    (fmt(12, 0, False), fmt(14, 2), fmt(30, 6), fmt(20, 4), 6, 3, "FOLD_EVEN"),   # unsigned samples: -x wraps back into IN_TYPE
    (_Q15, _Q15, fmt(18, 1), fmt(18, 1), 8, 2, "FOLD_EVEN"),                    # accumulator too narrow: wraps
    # order-dependent accumulators / saturating sample negation
    (fmt(16, 1, True, TRN, "AC_SAT"), _Q15, fmt(24, 4, True, TRN, "AC_SAT"), fmt(16, 1, True, "AC_RND_CONV", "AC_SAT_SYM"), 8, 4, "FOLD_EVEN"),
    (_Q15, _Q15, fmt(30, 6, True, "AC_TRN_ZERO", "AC_SAT_ZERO"), fmt(12, 1, True, "AC_RND_INF", "AC_SAT"), 5, 2, "FOLD_ODD"),
    (_Q15, _Q15, fmt(24, 4, True, TRN, "AC_SAT"), fmt(16, 1), 7, 3, "FOLD_ANTI"),
]


def pi_coeffsz(cfg):
    """COEFFSZ an instantiation reads: ac_poly_intr.h:141 (i + j*NTAPS/2), :199 (i + (NTAPS/2+1)*j), :249 (i + NTAPS*j)."""
    nt, IF, ft = cfg[4], cfg[5], cfg[6]
    return IF * (nt // 2 if ft == "FOLD_EVEN" else (nt // 2 + 1 if ft == "FOLD_ODD" else nt))


# ac_intg_dump instantiations: (in, acc, out, NS, CHN)
ID_CONFIGS = [
    (_Q15, fmt(32, 17), fmt(32, 17), 64, 4),
    (_Q15, fmt(40, 25), fmt(40, 25), 1024, 2),
    (fmt(32, 16), fmt(64, 32), fmt(64, 32), 1024, 4),               # the header's usage example (ac_intg_dump.h:46-51)
    (_Q15, fmt(24, 12), fmt(16, 8), 16, 3),                          # F_acc < F_in: every add truncates
    (_Q15, fmt(24, 12, True, RND), fmt(16, 8, True, RND), 16, 1),
    (fmt(12, 0, False), fmt(20, 8, False), fmt(20, 8, False), 100, 5),
    (_Q15, fmt(20, 5, True, TRN, "AC_SAT"), fmt(12, 2, True, "AC_RND_CONV", "AC_SAT_SYM"), 32, 2),   # order-dependent
    (_Q15, fmt(18, 3), fmt(18, 3), 32, 8),                            # accumulator wraps
]


# ac_mv_avg instantiations: (MAX_SAMPLE, TAPS, WIN_TYPE, in, out, acc, coeff).  Oracle A for these
# is the unmodified ac_mv_avg.h over a RESTATED ac_window_1d_flag (oracle/ac_shim/ac_window.h): parity unpinned.
MV_CONFIGS = [
    (1024, 7, "AC_CLIP", fmt(16, 2), fmt(16, 2), fmt(16, 2), fmt(16, 2)),                 # the manual's usage example (pdf p.30)
    (1024, 7, "AC_MIRROR", fmt(16, 2), fmt(16, 2), fmt(16, 2), fmt(16, 2)),
    (1024, 7, "AC_WIN", fmt(16, 2), fmt(16, 2), fmt(16, 2), fmt(16, 2)),
    (4096, 31, "AC_MIRROR", _Q15, _ACC40, _ACC40, _Q15),                                 # exact accumulation
    (4096, 31, "AC_CLIP", _Q15, fmt(20, 3, True, RND, "AC_SAT"), fmt(28, 6, True, RND), fmt(12, 1)),
    (500, 3, "AC_CLIP", fmt(12, 0, False), fmt(24, 8, False), fmt(24, 8, False), fmt(10, 2, False)),
    (2000, 9, "AC_MIRROR", fmt(24, 8), fmt(20, 6, True, "AC_RND_CONV", "AC_SAT_SYM"), fmt(30, 10, True, "AC_TRN_ZERO", "AC_SAT"), fmt(16, 1)),
    (300, 1, "AC_CLIP", _Q15, _ACC40, _ACC40, _Q15),                                      # TAPS = 1: a scaled copy
    (64, 5, "AC_WIN", fmt(32, 16), fmt(64, 32), fmt(64, 32), fmt(32, 16)),
    (8192, 63, "AC_CLIP", _Q15, _ACC40, fmt(18, 3), _Q15),                                 # the (ACC_TYPE) cast of the sample drops bits
]


def rs_ram_words(cfg):
    """Coefficient RAM words an instantiation reads (ac_fir_reg_share.h:122-133 and analogues)."""
    N, _fi, _fo, _fc, _fa, mww, bs, bo, ft = cfg
    used = N if ft == "SHIFT_REG" else (N // 2 if ft.startswith("FOLD_EVEN") else (N - 1) // 2 + 1)
    nblk = (used + bs - 1) // bs
    return (nblk - 1) * mww + bo + bs


def fir_configs():
    """Flat list of (id, name, in, coeff, acc, out, taps)."""
    res = []
    for name, fi, fc, fa, fo, taps in FIR_FORMATS:
        for t in taps:
            res.append((len(res), name, fi, fc, fa, fo, t))
    return res
