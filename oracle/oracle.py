"""oracle/oracle.py -- TEST INFRASTRUCTURE: ctypes loaders for the two CPU oracles.

* Oracle B  (`liboracle_b.so`, oracle_b.c): header-free integer restatement; always available.
* Oracle A  (`_ref/libacdsp_ref.so`, ref_driver_*.cpp): the UNMODIFIED reference templates
  compiled from /root/reference over the clean-room shim; compile-time configurations only
  (ref_configs.py).  Built in the dev container; the .so travels to the GPU box.

All values are raw two's-complement integers in numpy int64 arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import ref_configs as rc

HERE = os.path.dirname(os.path.abspath(__file__))
Q_MODES = ["AC_TRN", "AC_RND", "AC_TRN_ZERO", "AC_RND_ZERO", "AC_RND_INF", "AC_RND_MIN_INF", "AC_RND_CONV", "AC_RND_CONV_ODD"]
O_MODES = ["AC_WRAP", "AC_SAT", "AC_SAT_ZERO", "AC_SAT_SYM"]
FTYPES = ["SHIFT_REG", "ROTATE_SHIFT", "C_BUFF", "FOLD_EVEN", "FOLD_ODD", "TRANSPOSED", "FOLD_EVEN_ANTI", "FOLD_ODD_ANTI"]
FIR_CLASSES = ["const", "load", "prog"]


class ObFmt(C.Structure):
    _fields_ = [("W", C.c_int), ("I", C.c_int), ("S", C.c_int), ("Q", C.c_int), ("O", C.c_int)]


def normfmt(f):
    """(W, I[, S[, Q[, O]]]) with Q/O as names or ints -> (W, I, bool S, Qname, Oname)."""
    f = tuple(f)
    W, I = int(f[0]), int(f[1])
    S = bool(f[2]) if len(f) > 2 else True
    Q = f[3] if len(f) > 3 else "AC_TRN"
    O = f[4] if len(f) > 4 else "AC_WRAP"
    if not isinstance(Q, str):
        Q = Q_MODES[int(Q)]
    if not isinstance(O, str):
        O = O_MODES[int(O)]
    return (W, I, S, Q, O)


def _obfmt(f):
    W, I, S, Q, O = normfmt(f)
    return ObFmt(W, I, int(S), Q_MODES.index(Q), O_MODES.index(O))


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def build(force=False):
    """Compile Oracle B (and Oracle A when /root/reference is present). Building the checker is not using it."""
    ref_here = os.path.exists(os.environ.get("AC_DSP_REF", "/root/reference") + "/include/ac_dsp/ac_fir_load_coeffs.h")
    # with the reference tree at hand (dev container) make decides what is stale (shim, drivers, facade benches vs the
    # engine library); on the GPU box only the plain-C restatement can be (re)built and the prebuilt oracle/_ref is used
    if force or ref_here or not os.path.exists(os.path.join(HERE, "liboracle_b.so")):
        import fcntl
        with open(os.path.join(HERE, ".build.lock"), "w") as lock:          # several ranks / test workers may get here at once
            fcntl.flock(lock, fcntl.LOCK_EX)
            subprocess.check_call(["make", "-s", "-C", HERE, "-j8"] + (["-B"] if force else []), stdout=subprocess.DEVNULL)


_lib_b = None
_lib_a = None


def _libpath(name, sub):
    """B2D_ORACLE_LIBDIR points the loader at an alternative build of the two oracle libraries (the UBSan build that
    `make -C oracle sanitize` leaves in oracle/_ref/san); default: the normal build."""
    alt = os.environ.get("B2D_ORACLE_LIBDIR")
    return os.path.join(alt, name) if alt else os.path.join(HERE, sub, name)


def lib_b():
    global _lib_b
    if _lib_b is None:
        path = _libpath("liboracle_b.so", "")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.ob_fir_create.restype = C.c_void_p
        L.ob_fir_create.argtypes = [C.POINTER(ObFmt)] * 4 + [C.c_int, C.c_int]
        L.ob_fir_destroy.argtypes = [C.c_void_p]
        L.ob_fir_load.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.ob_fir_run.restype = C.c_long
        L.ob_fir_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64)]
        L.ob_cic_create.restype = C.c_void_p
        L.ob_cic_create.argtypes = [C.c_int, C.POINTER(ObFmt), C.POINTER(ObFmt), C.c_int, C.c_int, C.c_int]
        L.ob_cic_destroy.argtypes = [C.c_void_p]
        L.ob_cic_run.restype = C.c_long
        L.ob_cic_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64)]
        L.ob_cic_int_width.restype = C.c_int
        L.ob_cic_int_width.argtypes = [C.c_int, C.POINTER(ObFmt), C.c_int, C.c_int, C.c_int]
        L.ob_pd_create.restype = C.c_void_p
        L.ob_pd_create.argtypes = [C.POINTER(ObFmt)] * 4 + [C.c_int, C.c_int]
        L.ob_pd_destroy.argtypes = [C.c_void_p]
        L.ob_pd_load.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.ob_pd_run.restype = C.c_long
        L.ob_pd_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64)]
        L.ob_pi_create.restype = C.c_void_p
        L.ob_pi_create.argtypes = [C.POINTER(ObFmt)] * 4 + [C.c_int, C.c_int, C.c_int]
        L.ob_pi_destroy.argtypes = [C.c_void_p]
        L.ob_pi_load.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 3
        L.ob_pi_run.restype = C.c_long
        L.ob_pi_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64)]
        L.ob_id_create.restype = C.c_void_p
        L.ob_id_create.argtypes = [C.POINTER(ObFmt)] * 3 + [C.c_int, C.c_int]
        L.ob_id_destroy.argtypes = [C.c_void_p]
        L.ob_id_run.restype = C.c_long
        L.ob_id_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64)]
        L.ob_rs_create.restype = C.c_void_p
        L.ob_rs_create.argtypes = [C.POINTER(ObFmt)] * 4 + [C.c_int] * 5
        L.ob_rs_destroy.argtypes = [C.c_void_p]
        L.ob_rs_ram_words.argtypes = [C.c_void_p]
        L.ob_rs_run.restype = C.c_long
        L.ob_rs_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_int64)]
        L.ob_rs_delay_out.restype = C.c_int64
        L.ob_rs_delay_out.argtypes = [C.c_void_p]
        _lib_b = L
    return _lib_b


def have_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "libacdsp_ref.so"))


def lib_a():
    global _lib_a
    if _lib_a is None:
        L = C.CDLL(_libpath("libacdsp_ref.so", "_ref"))
        L.acref_fir_create.restype = C.c_void_p
        L.acref_fir_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.acref_fir_load.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.acref_fir_run.restype = C.c_long
        L.acref_fir_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64)]
        L.acref_fir_destroy.argtypes = [C.c_void_p]
        L.acref_fir_last_seconds.restype = C.c_double
        L.acref_fir_last_seconds.argtypes = [C.c_void_p]
        L.acref_cic_last_seconds.restype = C.c_double
        L.acref_cic_last_seconds.argtypes = [C.c_void_p]
        for fn in (L.acref_cic_dec_create, L.acref_cic_intr_create):
            fn.restype = C.c_void_p
            fn.argtypes = [C.c_int]
        L.acref_cic_run.restype = C.c_long
        L.acref_cic_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64)]
        L.acref_cic_destroy.argtypes = [C.c_void_p]
        L.acref_pd_create.restype = C.c_void_p
        L.acref_pd_create.argtypes = [C.c_int]
        L.acref_pd_load.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.acref_pd_run.restype = C.c_long
        L.acref_pd_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64)]
        L.acref_pd_destroy.argtypes = [C.c_void_p]
        L.acref_pi_create.restype = C.c_void_p
        L.acref_pi_create.argtypes = [C.c_int]
        L.acref_pi_load.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 3
        L.acref_pi_run.restype = C.c_long
        L.acref_pi_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64)]
        L.acref_pi_destroy.argtypes = [C.c_void_p]
        L.acref_id_create.restype = C.c_void_p
        L.acref_id_create.argtypes = [C.c_int]
        L.acref_id_run.restype = C.c_long
        L.acref_id_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64)]
        L.acref_id_destroy.argtypes = [C.c_void_p]
        L.acref_rs_create.restype = C.c_void_p
        L.acref_rs_create.argtypes = [C.c_int]
        L.acref_rs_run.restype = C.c_long
        L.acref_rs_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.acref_rs_delay_out.restype = C.c_int64
        L.acref_rs_delay_out.argtypes = [C.c_void_p]
        L.acref_rs_ram_words.argtypes = [C.c_void_p]
        L.acref_rs_destroy.argtypes = [C.c_void_p]
        _lib_a = L
    return _lib_a


# --------------------------------------------------------------------------- Oracle B objects
class FirB:
    """Oracle B FIR object: load(coeffs) then run(samples) any number of times (state persists)."""

    def __init__(self, fin, fcoeff, facc, fout, n_taps, ftype):
        self.L = lib_b()
        ft = FTYPES.index(ftype) if isinstance(ftype, str) else int(ftype)
        a, b, c, d = _obfmt(fin), _obfmt(fcoeff), _obfmt(facc), _obfmt(fout)
        self.h = self.L.ob_fir_create(C.byref(a), C.byref(b), C.byref(c), C.byref(d), int(n_taps), ft)
        self.n_taps = int(n_taps)

    def load(self, coeffs):
        c = _i64(coeffs)
        assert c.size == self.n_taps
        self.L.ob_fir_load(self.h, _p(c))

    def run(self, x):
        x = _i64(x)
        out = np.empty(x.size, dtype=np.int64)
        n = self.L.ob_fir_run(self.h, _p(x), x.size, _p(out))
        if n < 0:
            raise ValueError("unsupported ftype")
        return out[:n]

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ob_fir_destroy(self.h)
            self.h = None


class CicB:
    def __init__(self, mode, fin, fout, R, M, N):
        self.L = lib_b()
        self.intr = 1 if mode == "intr" else 0
        a, b = _obfmt(fin), _obfmt(fout)
        self.h = self.L.ob_cic_create(self.intr, C.byref(a), C.byref(b), int(R), int(M), int(N))
        self.R, self.M, self.N = int(R), int(M), int(N)

    def run(self, x):
        x = _i64(x)
        cap = x.size * self.R + self.R + 8 if self.intr else x.size // self.R + 2
        out = np.empty(max(cap, 1), dtype=np.int64)
        n = self.L.ob_cic_run(self.h, _p(x), x.size, _p(out))
        return out[:n].copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ob_cic_destroy(self.h)
            self.h = None


def cic_int_width(mode, fin, R, M, N):
    a = _obfmt(fin)
    return lib_b().ob_cic_int_width(1 if mode == "intr" else 0, C.byref(a), R, M, N)


# --------------------------------------------------------------------------- Oracle A objects
def ref_fir_cfg_id(fin, fcoeff, facc, fout, n_taps):
    key = (normfmt(fin), normfmt(fcoeff), normfmt(facc), normfmt(fout), int(n_taps))
    for cid, _name, fi, fc, fa, fo, t in rc.fir_configs():
        if (fi, fc, fa, fo, t) == key:
            return cid
    return None


def ref_cic_cfg_id(mode, fin, fout, R, M, N):
    key = (mode, int(R), int(M), int(N), normfmt(fin), normfmt(fout))
    for cid, c in enumerate(rc.CIC_CONFIGS):
        if c == key:
            return cid
    return None


class FirA:
    """The real reference class (const / load / prog) for one compiled-in configuration."""

    def __init__(self, cls, fin, fcoeff, facc, fout, n_taps, ftype):
        self.L = lib_a()
        cid = ref_fir_cfg_id(fin, fcoeff, facc, fout, n_taps)
        if cid is None:
            raise KeyError("configuration not instantiated in oracle/_ref (add it to ref_configs.py)")
        ft = FTYPES.index(ftype) if isinstance(ftype, str) else int(ftype)
        k = FIR_CLASSES.index(cls) if isinstance(cls, str) else int(cls)
        self.h = self.L.acref_fir_create(cid, k, ft)
        if not self.h:
            raise ValueError("reference does not dispatch this ftype")
        self.n_taps = int(n_taps)

    def load(self, coeffs):
        c = _i64(coeffs)
        assert c.size == self.n_taps
        if self.L.acref_fir_load(self.h, _p(c)) != 0:
            raise RuntimeError("constant-coefficient filter: coefficients are fixed at construction")

    def run(self, x):
        x = _i64(x)
        out = np.empty(x.size, dtype=np.int64)
        n = self.L.acref_fir_run(self.h, _p(x), x.size, _p(out))
        if n < 0:
            raise RuntimeError("coefficients not set")
        return out[:n]

    def last_run_seconds(self):
        """Seconds spent inside the reference's run() during the last run() (channel fill / drain excluded)."""
        return self.L.acref_fir_last_seconds(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.acref_fir_destroy(self.h)
            self.h = None


class CicA:
    def __init__(self, mode, fin, fout, R, M, N):
        self.L = lib_a()
        cid = ref_cic_cfg_id(mode, fin, fout, R, M, N)
        if cid is None:
            raise KeyError("configuration not instantiated in oracle/_ref (add it to ref_configs.py)")
        self.h = (self.L.acref_cic_intr_create if mode == "intr" else self.L.acref_cic_dec_create)(cid)
        self.intr = mode == "intr"
        self.R = int(R)

    def run(self, x):
        x = _i64(x)
        cap = x.size * self.R + self.R + 8 if self.intr else x.size // self.R + 2
        out = np.empty(max(cap, 1), dtype=np.int64)
        n = self.L.acref_cic_run(self.h, _p(x), x.size, _p(out))
        return out[:n].copy()

    def last_run_seconds(self):
        return self.L.acref_cic_last_seconds(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.acref_cic_destroy(self.h)
            self.h = None


# --------------------------------------------------------------------------- ac_fir_reg_share (row N1)
class RsB:
    """Oracle B ac_fir_reg_share object: run(samples, coefficient RAM) = one reference run() per sample."""

    def __init__(self, fin, fout, fcoeff, facc, n_taps, mww=1, blk_sz=1, blk_off=0, ftype="SHIFT_REG"):
        self.L = lib_b()
        ft = FTYPES.index(ftype) if isinstance(ftype, str) else int(ftype)
        a, b, c, d = _obfmt(fin), _obfmt(fcoeff), _obfmt(facc), _obfmt(fout)
        self.h = self.L.ob_rs_create(C.byref(a), C.byref(b), C.byref(c), C.byref(d), int(n_taps), int(mww), int(blk_sz), int(blk_off), ft)
        self.ram_words = self.L.ob_rs_ram_words(self.h)

    def run(self, x, ram):
        x, ram = _i64(x), _i64(ram)
        out = np.empty(x.size, dtype=np.int64)
        n = self.L.ob_rs_run(self.h, _p(x), x.size, _p(ram), ram.size, _p(out))
        if n < 0:
            raise ValueError("unsupported ftype / coefficient RAM too small")
        return out[:n]

    def delay_out(self):
        return int(self.L.ob_rs_delay_out(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ob_rs_destroy(self.h)
            self.h = None


class RsA:
    """The real reference ac_fir_reg_share for one compiled-in configuration (index into ref_configs.RS_CONFIGS)."""

    def __init__(self, cfg_id):
        self.L = lib_a()
        self.h = self.L.acref_rs_create(int(cfg_id))
        if not self.h:
            raise KeyError("configuration not instantiated in oracle/_ref")
        self.ram_words = self.L.acref_rs_ram_words(self.h)

    def run(self, x, ram):
        x, ram = _i64(x), _i64(ram)
        assert ram.size >= self.ram_words
        out = np.empty(x.size, dtype=np.int64)
        n = self.L.acref_rs_run(self.h, _p(x), x.size, _p(ram), _p(out))
        return out[:n]

    def delay_out(self):
        return int(self.L.acref_rs_delay_out(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.acref_rs_destroy(self.h)
            self.h = None


# --------------------------------------------------------------------------- ac_poly_dec (row N2)
class PdB:
    """Oracle B ac_poly_dec: load(coeffs[NTAPS*DF], phase order) then run(samples) -> floor((pending + n) / DF) outputs."""

    def __init__(self, fin, fcoeff, facc, fout, ntaps, df):
        self.L = lib_b()
        a, b, c, d = _obfmt(fin), _obfmt(fcoeff), _obfmt(facc), _obfmt(fout)
        self.h = self.L.ob_pd_create(C.byref(a), C.byref(b), C.byref(c), C.byref(d), int(ntaps), int(df))
        self.n_coeff, self.df = int(ntaps) * int(df), int(df)

    def load(self, coeffs):
        c = _i64(coeffs)
        assert c.size == self.n_coeff
        self.L.ob_pd_load(self.h, _p(c))

    def run(self, x):
        x = _i64(x)
        out = np.empty(x.size // self.df + 2, dtype=np.int64)
        n = self.L.ob_pd_run(self.h, _p(x), x.size, _p(out))
        return out[:n].copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ob_pd_destroy(self.h)
            self.h = None


class PdA:
    """The real reference ac_poly_dec for one compiled-in configuration (index into ref_configs.PD_CONFIGS)."""

    def __init__(self, cfg_id):
        self.L = lib_a()
        self.h = self.L.acref_pd_create(int(cfg_id))
        if not self.h:
            raise KeyError("configuration not instantiated in oracle/_ref")
        _fi, _fc, _fa, _fo, nt, df = rc.PD_CONFIGS[cfg_id]
        self.n_coeff, self.df = nt * df, df

    def load(self, coeffs):
        c = _i64(coeffs)
        assert c.size == self.n_coeff
        self.L.acref_pd_load(self.h, _p(c))

    def run(self, x):
        x = _i64(x)
        out = np.empty(x.size // self.df + 2, dtype=np.int64)
        n = self.L.acref_pd_run(self.h, _p(x), x.size, _p(out))
        return out[:n].copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.acref_pd_destroy(self.h)
            self.h = None


# --------------------------------------------------------------------------- ac_intg_dump (row N4)
def id_frame_samples(n_sample, NS, CHN):
    """Samples (all channels) a frame with token n_sample consumes: n_sample * CHN if it dumps, else NS * CHN."""
    return (int(n_sample) if 1 <= int(n_sample) <= NS else NS) * CHN


class PiB:
    """Oracle B ac_poly_intr: load(coeffs, sign, corr); run(samples) -> IF outputs per step (folded forms: one step late)."""

    def __init__(self, fin, fcoeff, facc, fout, ntaps, IF, ftype):
        self.L = lib_b()
        a, b, c, d = _obfmt(fin), _obfmt(fcoeff), _obfmt(facc), _obfmt(fout)
        self.nt, self.IF, self.ftype = int(ntaps), int(IF), ftype
        self.h = self.L.ob_pi_create(C.byref(a), C.byref(b), C.byref(c), C.byref(d), self.nt, self.IF, rc.PI_FTYPES.index(ftype))
        self.coeffsz = rc.pi_coeffsz((None, None, None, None, self.nt, self.IF, ftype))

    def load(self, coeffs, sign=None, corr=None):
        c = _i64(coeffs)
        assert c.size == self.coeffsz
        sg = _i64(np.ones(self.IF) if sign is None else sign)
        cr = _i64(np.arange(self.IF) if corr is None else corr)
        self.L.ob_pi_load(self.h, _p(c), _p(sg), _p(cr))

    def run(self, x):
        x = _i64(x)
        out = np.empty(x.size * self.IF + 1, dtype=np.int64)
        n = self.L.ob_pi_run(self.h, _p(x), x.size, _p(out))
        return out[:n].copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ob_pi_destroy(self.h)
            self.h = None


class PiA:
    """The real reference ac_poly_intr for one compiled-in configuration (index into ref_configs.PI_CONFIGS)."""

    def __init__(self, cfg_id):
        self.L = lib_a()
        self.h = self.L.acref_pi_create(int(cfg_id))
        if not self.h:
            raise KeyError("configuration not instantiated in oracle/_ref")
        cfg = rc.PI_CONFIGS[cfg_id]
        self.nt, self.IF, self.ftype = cfg[4], cfg[5], cfg[6]
        self.coeffsz = rc.pi_coeffsz(cfg)

    def load(self, coeffs, sign=None, corr=None):
        c = _i64(coeffs)
        assert c.size == self.coeffsz
        sg = _i64(np.ones(self.IF) if sign is None else sign)
        cr = _i64(np.arange(self.IF) if corr is None else corr)
        self.L.acref_pi_load(self.h, _p(c), _p(sg), _p(cr))

    def run(self, x):
        x = _i64(x)
        out = np.empty(x.size * self.IF + 1, dtype=np.int64)
        n = self.L.acref_pi_run(self.h, _p(x), x.size, _p(out))
        return out[:n].copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.acref_pi_destroy(self.h)
            self.h = None


class IdB:
    """Oracle B ac_intg_dump: run(samples interleaved over CHN, n_sample tokens) -> CHN outputs per dumping frame."""

    def __init__(self, fin, facc, fout, NS, CHN):
        self.L = lib_b()
        a, b, c = _obfmt(fin), _obfmt(facc), _obfmt(fout)
        self.h = self.L.ob_id_create(C.byref(a), C.byref(b), C.byref(c), int(NS), int(CHN))
        self.NS, self.CHN = int(NS), int(CHN)

    def run(self, x, n_sample):
        x, ns = _i64(x), _i64(n_sample)
        out = np.empty(ns.size * self.CHN + 1, dtype=np.int64)
        n = self.L.ob_id_run(self.h, _p(x), x.size, _p(ns), ns.size, _p(out))
        if n < 0:
            raise ValueError("samples do not cover the frames exactly")
        return out[:n].copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ob_id_destroy(self.h)
            self.h = None


WIN_MODES = ["AC_WIN", "AC_CLIP", "AC_MIRROR"]


def mv_run_b(fin, fout, facc, fcoeff, taps, win, coeffs, x, n_sample):
    """Oracle B ac_mv_avg (parity unpinned: ac_window restated): one run() call over whole bursts of n_sample samples."""
    L = lib_b()
    L.ob_mvavg_run.restype = C.c_long
    L.ob_mvavg_run.argtypes = [C.POINTER(ObFmt)] * 4 + [C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_long, C.c_long,
                               C.POINTER(C.c_int64)]
    a, b, c, d = _obfmt(fin), _obfmt(fout), _obfmt(facc), _obfmt(fcoeff)
    x, h = _i64(x), _i64(coeffs)
    out = np.empty(x.size + 1, dtype=np.int64)
    w = WIN_MODES.index(win) if isinstance(win, str) else int(win)
    n = L.ob_mvavg_run(C.byref(a), C.byref(b), C.byref(c), C.byref(d), int(taps), w, _p(h), _p(x), x.size, int(n_sample), _p(out))
    if n < 0:
        raise ValueError("bursts must be whole and at least TAPS samples long")
    return out[:n].copy()


class MvA:
    """The unmodified reference ac_mv_avg for one compiled-in configuration (index into ref_configs.MV_CONFIGS), over the
    restated ac_window_1d_flag of oracle/ac_shim/ac_window.h."""

    def __init__(self, cfg_id, coeffs):
        self.L = lib_a()
        self.L.acref_mv_create.restype = C.c_void_p
        self.L.acref_mv_create.argtypes = [C.c_int, C.POINTER(C.c_int64)]
        self.L.acref_mv_run.restype = C.c_long
        self.L.acref_mv_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.c_longlong, C.POINTER(C.c_int64)]
        self.L.acref_mv_destroy.argtypes = [C.c_void_p]
        h = _i64(coeffs)
        self.h = self.L.acref_mv_create(int(cfg_id), _p(h))
        if not self.h:
            raise KeyError("configuration not instantiated in oracle/_ref")

    def run(self, x, n_sample):
        x = _i64(x)
        out = np.empty(x.size + 1, dtype=np.int64)
        n = self.L.acref_mv_run(self.h, _p(x), x.size, int(n_sample), _p(out))
        return out[:n].copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.acref_mv_destroy(self.h)
            self.h = None


class IdA:
    """The real reference ac_intg_dump for one compiled-in configuration (index into ref_configs.ID_CONFIGS)."""

    def __init__(self, cfg_id):
        self.L = lib_a()
        self.h = self.L.acref_id_create(int(cfg_id))
        if not self.h:
            raise KeyError("configuration not instantiated in oracle/_ref")
        _fi, _fa, _fo, self.NS, self.CHN = rc.ID_CONFIGS[cfg_id]

    def run(self, x, n_sample):
        x, ns = _i64(x), _i64(n_sample)
        assert x.size == sum(id_frame_samples(v, self.NS, self.CHN) for v in ns), "the reference reads past the end of its channel otherwise"
        out = np.empty(ns.size * self.CHN + 1, dtype=np.int64)
        n = self.L.acref_id_run(self.h, _p(x), x.size, _p(ns), ns.size, _p(out))
        return out[:n].copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.acref_id_destroy(self.h)
            self.h = None


# --------------------------------------------------------------------------- helpers
def rand_raw(rng, f, n, kind="uniform"):
    """Full-range random raw values for format f."""
    W, _I, S, _Q, _O = normfmt(f)
    lo, hi = (-(1 << (W - 1)), (1 << (W - 1)) - 1) if S else (0, (1 << W) - 1)
    if kind == "min":
        return np.full(n, lo, dtype=np.int64)
    if kind == "max":
        return np.full(n, hi, dtype=np.int64)
    if kind == "alt":
        a = np.full(n, hi, dtype=np.int64)
        a[1::2] = lo
        return a
    return rng.integers(lo, hi, size=n, endpoint=True, dtype=np.int64)
