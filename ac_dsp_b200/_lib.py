"""ctypes binding of libb200dsp.so -- the C-ABI declared in include/b200dsp.h.

Loading fails loudly when the CUDA library is absent; nothing here computes on the CPU.
"""
import ctypes as C
import os

from . import build as _build

Q_MODES = ["AC_TRN", "AC_RND", "AC_TRN_ZERO", "AC_RND_ZERO", "AC_RND_INF", "AC_RND_MIN_INF", "AC_RND_CONV", "AC_RND_CONV_ODD"]
O_MODES = ["AC_WRAP", "AC_SAT", "AC_SAT_ZERO", "AC_SAT_SYM"]
FTYPES = ["SHIFT_REG", "ROTATE_SHIFT", "C_BUFF", "FOLD_EVEN", "FOLD_ODD", "TRANSPOSED", "FOLD_EVEN_ANTI", "FOLD_ODD_ANTI"]
FIR_KINDS = ["const", "load", "prog", "reg_share"]
PLANAR, INTERLEAVED = 0, 1
WIRE_CONTAINER, WIRE_PACKED = 0, 1

OK, EUNSUPPORTED, EINVAL, ECUDA, ENCCL, ENOMEM, ESTATE = 0, -1, -2, -3, -4, -5, -6


class B2dFmt(C.Structure):
    _fields_ = [("W", C.c_int32), ("I", C.c_int32), ("S", C.c_int32), ("Q", C.c_int32), ("O", C.c_int32)]


class B2dFirDesc(C.Structure):
    _fields_ = [("fin", B2dFmt), ("coeff", B2dFmt), ("acc", B2dFmt), ("out", B2dFmt),
                ("n_taps", C.c_uint32), ("ftype", C.c_int32), ("kind", C.c_int32),
                ("n_channels", C.c_uint32), ("layout", C.c_int32), ("device", C.c_int32)]


class B2dCicDesc(C.Structure):
    _fields_ = [("fin", B2dFmt), ("out", B2dFmt), ("R", C.c_uint32), ("M", C.c_uint32), ("N", C.c_uint32),
                ("mode", C.c_int32), ("n_channels", C.c_uint32), ("layout", C.c_int32), ("device", C.c_int32)]


class B2dPolydecDesc(C.Structure):
    _fields_ = [("fin", B2dFmt), ("coeff", B2dFmt), ("acc", B2dFmt), ("out", B2dFmt), ("n_taps", C.c_uint32), ("df", C.c_uint32),
                ("n_channels", C.c_uint32), ("layout", C.c_int32), ("device", C.c_int32)]


class B2dPolyintrDesc(C.Structure):
    _fields_ = [("fin", B2dFmt), ("coeff", B2dFmt), ("acc", B2dFmt), ("out", B2dFmt), ("n_taps", C.c_uint32), ("intr_factor", C.c_uint32),
                ("ftype", C.c_int32), ("n_channels", C.c_uint32), ("layout", C.c_int32), ("device", C.c_int32)]


PI_FTYPES = ["FOLD_EVEN", "FOLD_ODD", "FOLD_ANTI"]   # the polyphase enum of ac_poly_intr.h:95


class B2dIntgdumpDesc(C.Structure):
    _fields_ = [("fin", B2dFmt), ("acc", B2dFmt), ("out", B2dFmt), ("ns", C.c_uint32), ("chn", C.c_uint32), ("device", C.c_int32)]


class B2dMvavgDesc(C.Structure):
    _fields_ = [("fin", B2dFmt), ("out", B2dFmt), ("acc", B2dFmt), ("coeff", B2dFmt), ("max_sample", C.c_uint32), ("taps", C.c_uint32),
                ("win_type", C.c_int32), ("device", C.c_int32)]


WIN_MODES = ["AC_WIN", "AC_CLIP", "AC_MIRROR"]   # b2d_window_mode


class B2dError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(msg)
        self.status = status


_lib = None


def lib_path():
    """libb200dsp.so of this tree; B2D_LIBRARY points at another build of it (kernel A/B runs: tools/build_variant.sh)."""
    return os.environ.get("B2D_LIBRARY") or _build.LIB


def load():
    """The loaded library (ctypes.CDLL). Raises if libb200dsp.so is missing -- there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `python -m ac_dsp_b200.build` (needs nvcc); "
                          "the engine has no CPU fallback")
    L = C.CDLL(path)
    vp, sz, i32, u32 = C.c_void_p, C.c_size_t, C.c_int32, C.c_uint32
    psz = C.POINTER(C.c_size_t)
    sig = {
        "b2d_version": (C.c_char_p, []), "b2d_strerror": (C.c_char_p, [C.c_int]), "b2d_last_error": (C.c_char_p, []),
        "b2d_container_bytes": (C.c_int, [i32]), "b2d_device_count": (C.c_int, []),
        "b2d_host_alloc": (C.c_int, [C.POINTER(vp), sz]), "b2d_host_free": (C.c_int, [vp]),
        "b2d_fir_create": (C.c_int, [C.POINTER(vp), C.POINTER(B2dFirDesc)]), "b2d_fir_destroy": (C.c_int, [vp]),
        "b2d_fir_load": (C.c_int, [vp, vp, sz, i32]), "b2d_fir_set_comm": (C.c_int, [vp, vp, i32]),
        "b2d_fir_run": (C.c_int, [vp, vp, sz, vp, psz]), "b2d_fir_run_dev": (C.c_int, [vp, vp, sz, vp, psz, vp]),
        "b2d_fir_reset": (C.c_int, [vp]), "b2d_fir_state_bytes": (C.c_int, [vp, psz]),
        "b2d_fir_get_state": (C.c_int, [vp, vp, sz]), "b2d_fir_set_state": (C.c_int, [vp, vp, sz]),
        "b2d_fir_path": (C.c_char_p, [vp]),
        "b2d_fir_ovs_margin": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "b2d_fir_load_blocked": (C.c_int, [vp, vp, sz, u32, u32, u32, i32]), "b2d_fir_delay_line_out": (C.c_int, [vp, vp]),
        "b2d_fir_run_window": (C.c_int, [vp, vp, vp]),
        "b2d_cic_create": (C.c_int, [C.POINTER(vp), C.POINTER(B2dCicDesc)]), "b2d_cic_destroy": (C.c_int, [vp]),
        "b2d_cic_int_width": (C.c_int, [C.POINTER(B2dCicDesc), C.POINTER(i32)]), "b2d_cic_max_out": (sz, [vp, sz]),
        "b2d_cic_run": (C.c_int, [vp, vp, sz, vp, psz]), "b2d_cic_run_dev": (C.c_int, [vp, vp, sz, vp, psz, vp]),
        "b2d_cic_reset": (C.c_int, [vp]), "b2d_cic_state_bytes": (C.c_int, [vp, psz]),
        "b2d_cic_get_state": (C.c_int, [vp, vp, sz]), "b2d_cic_set_state": (C.c_int, [vp, vp, sz]),
        "b2d_cic_path": (C.c_char_p, [vp]),
        "b2d_cicfir_create": (C.c_int, [C.POINTER(vp), C.POINTER(B2dCicDesc), C.POINTER(B2dFirDesc)]),
        "b2d_cicfir_destroy": (C.c_int, [vp]), "b2d_cicfir_load": (C.c_int, [vp, vp, sz, i32]),
        "b2d_cicfir_max_out": (sz, [vp, sz]), "b2d_cicfir_run": (C.c_int, [vp, vp, sz, vp, psz]),
        "b2d_cicfir_run_dev": (C.c_int, [vp, vp, sz, vp, psz, vp]), "b2d_cicfir_reset": (C.c_int, [vp]),
        "b2d_cicfir_path": (C.c_char_p, [vp]),
        "b2d_polydec_create": (C.c_int, [C.POINTER(vp), C.POINTER(B2dPolydecDesc)]), "b2d_polydec_destroy": (C.c_int, [vp]),
        "b2d_polydec_load": (C.c_int, [vp, vp, sz, i32]), "b2d_polydec_max_out": (sz, [vp, sz]),
        "b2d_polydec_run": (C.c_int, [vp, vp, sz, vp, psz]), "b2d_polydec_run_dev": (C.c_int, [vp, vp, sz, vp, psz, vp]),
        "b2d_polydec_reset": (C.c_int, [vp]), "b2d_polydec_path": (C.c_char_p, [vp]),
        "b2d_polyintr_create": (C.c_int, [C.POINTER(vp), C.POINTER(B2dPolyintrDesc)]), "b2d_polyintr_destroy": (C.c_int, [vp]),
        "b2d_polyintr_coeffsz": (sz, [vp]), "b2d_polyintr_load": (C.c_int, [vp, vp, sz, vp, vp, i32]), "b2d_polyintr_max_out": (sz, [vp, sz]),
        "b2d_polyintr_run": (C.c_int, [vp, vp, sz, vp, psz]), "b2d_polyintr_run_dev": (C.c_int, [vp, vp, sz, vp, psz, vp]),
        "b2d_polyintr_reset": (C.c_int, [vp]), "b2d_polyintr_path": (C.c_char_p, [vp]),
        "b2d_intgdump_create": (C.c_int, [C.POINTER(vp), C.POINTER(B2dIntgdumpDesc)]), "b2d_intgdump_destroy": (C.c_int, [vp]),
        "b2d_intgdump_run": (C.c_int, [vp, vp, sz, vp, sz, vp, psz]), "b2d_intgdump_run_dev": (C.c_int, [vp, vp, sz, vp, sz, vp, psz, vp]),
        "b2d_intgdump_reset": (C.c_int, [vp]), "b2d_intgdump_path": (C.c_char_p, [vp]),
        "b2d_cicfir_state_bytes": (C.c_int, [vp, psz]), "b2d_cicfir_get_state": (C.c_int, [vp, vp, sz]), "b2d_cicfir_set_state": (C.c_int, [vp, vp, sz]),
        "b2d_polydec_state_bytes": (C.c_int, [vp, psz]), "b2d_polydec_get_state": (C.c_int, [vp, vp, sz]), "b2d_polydec_set_state": (C.c_int, [vp, vp, sz]),
        "b2d_polyintr_state_bytes": (C.c_int, [vp, psz]), "b2d_polyintr_get_state": (C.c_int, [vp, vp, sz]), "b2d_polyintr_set_state": (C.c_int, [vp, vp, sz]),
        "b2d_intgdump_state_bytes": (C.c_int, [vp, psz]), "b2d_intgdump_get_state": (C.c_int, [vp, vp, sz]), "b2d_intgdump_set_state": (C.c_int, [vp, vp, sz]),
        "b2d_shard_count": (C.c_int, [u32, i32, i32, C.POINTER(u32)]), "b2d_comm_unique_id": (C.c_int, [vp]),
        "b2d_comm_create": (C.c_int, [C.POINTER(vp), vp, i32, i32, i32]), "b2d_comm_destroy": (C.c_int, [vp]),
        "b2d_comm_barrier": (C.c_int, [vp]),
        "b2d_mvavg_create": (C.c_int, [C.POINTER(vp), C.POINTER(B2dMvavgDesc), vp]), "b2d_mvavg_destroy": (C.c_int, [vp]),
        "b2d_mvavg_max_out": (sz, [vp, sz]), "b2d_mvavg_run": (C.c_int, [vp, vp, sz, sz, vp, psz]),
        "b2d_mvavg_run_dev": (C.c_int, [vp, vp, sz, sz, vp, psz, vp]), "b2d_mvavg_path": (C.c_char_p, [vp]),
        "b2d_wire_bytes": (C.c_int, [i32, i32]), "b2d_unpack_wire": (C.c_int, [vp, sz, i32, i32, vp]),
        "b2d_fir_set_wire": (C.c_int, [vp, i32]), "b2d_cic_set_wire": (C.c_int, [vp, i32]), "b2d_cicfir_set_wire": (C.c_int, [vp, i32]),
        "b2d_polydec_set_wire": (C.c_int, [vp, i32]), "b2d_polyintr_set_wire": (C.c_int, [vp, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(status):
    if status != OK:
        L = load()
        raise B2dError(status, f"{L.b2d_strerror(status).decode()}: {L.b2d_last_error().decode()}")


def make_fmt(f):
    """ac_fixed-like spec -> B2dFmt.  Accepts (W, I[, S[, Q[, O]]]) with Q/O as names or ints."""
    if isinstance(f, B2dFmt):
        return f
    f = tuple(f)
    W, I = int(f[0]), int(f[1])
    S = int(bool(f[2])) if len(f) > 2 else 1
    Q = f[3] if len(f) > 3 else 0
    O = f[4] if len(f) > 4 else 0
    if isinstance(Q, str):
        Q = Q_MODES.index(Q)
    if isinstance(O, str):
        O = O_MODES.index(O)
    return B2dFmt(W, I, S, int(Q), int(O))
