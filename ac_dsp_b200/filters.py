"""Host-side mirror of the reference's five hot-path class templates, on top of the C-ABI.

Same names, template parameters (as constructor arguments) and run()/load call surface as
  ac_fir_const_coeffs / ac_fir_load_coeffs / ac_fir_prog_coeffs   (reference include/ac_dsp/ac_fir_*_coeffs.h)
  ac_cic_dec_full / ac_cic_intr_full                               (reference include/ac_dsp/ac_cic_*_full.h)
`ac_channel<T>` FIFOs become arrays of RAW two's-complement integers: numpy arrays (host path, copies inside
the call) or torch CUDA tensors (device path, asynchronous on torch's current stream).  All arithmetic runs in
the CUDA library; this module only marshals.
"""
import ctypes as C

import numpy as np

from . import _lib as L

_NP_DT = {2: np.int16, 4: np.int32, 8: np.int64}


def ac_fixed(W, I, S=True, Q="AC_TRN", O="AC_WRAP"):
    """ac_fixed<W, I, S, Q, O> type descriptor."""
    return (int(W), int(I), bool(S), Q, O)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _container_dtype(fmt):
    return _NP_DT[L.load().b2d_container_bytes(fmt.W)]


def _host_dtype(fmt):
    """numpy dtype of host results: the container, read as unsigned for S = 0 formats (a full-width unsigned value,
    e.g. ac_fixed<32,.,false>, has its top bit set in the container; torch results stay in the signed container)."""
    dt = np.dtype(_container_dtype(fmt))
    return dt if fmt.S else np.dtype(f"u{dt.itemsize}")


def _get_state(h, prefix):
    lib = L.load()
    n = C.c_size_t(0)
    L.check(getattr(lib, f"b2d_{prefix}_state_bytes")(h, C.byref(n)))
    buf = np.zeros(n.value, dtype=np.uint8)
    L.check(getattr(lib, f"b2d_{prefix}_get_state")(h, buf.ctypes.data, buf.size))
    return buf


def _set_state(h, prefix, blob):
    blob = np.ascontiguousarray(np.asarray(blob, dtype=np.uint8))
    L.check(getattr(L.load(), f"b2d_{prefix}_set_state")(h, blob.ctypes.data, blob.size))


class _Block:
    """Shared marshaling: channel layout, numpy / torch dispatch."""

    _prefix = None      # C-ABI family prefix ("fir", "cic", ...) of the subclasses that have b2d_<prefix>_set_wire

    def set_wire(self, wire):
        """Format of the output array of the numpy (host-buffer) run(): "container" (default) or "packed" --
        ceil(W_out / 8) little-endian bytes per value (b200dsp.h: b2d_wire).  Packed runs return a uint8 array shaped
        like the container result with a trailing axis of that many bytes; unpack_wire() widens it on the host."""
        w = L.WIRE_PACKED if wire in ("packed", L.WIRE_PACKED) else L.WIRE_CONTAINER
        L.check(getattr(L.load(), f"b2d_{self._prefix}_set_wire")(self._h, w))
        self._wire = w

    def unpack_wire(self, packed):
        """Packed host output -> containers (b2d_unpack_wire), same shape without the trailing byte axis."""
        p = np.ascontiguousarray(packed, dtype=np.uint8)
        out = np.empty(p.shape[:-1], dtype=self._out_dt)
        L.check(L.load().b2d_unpack_wire(p.ctypes.data, out.size, self._out_W, self._out_S, out.ctypes.data))
        return out.view(self._host_dt)

    def _setup_io(self, fin, fout, n_channels, layout):
        self._wire = L.WIRE_CONTAINER
        self._out_W, self._out_S = int(fout.W), int(fout.S)
        self._C = int(n_channels)
        self._layout = L.INTERLEAVED if layout in ("interleaved", L.INTERLEAVED) else L.PLANAR
        self._in_dt = _container_dtype(fin)
        self._out_dt = _container_dtype(fout)
        self._host_dt = _host_dtype(fout)

    def _n_per_channel(self, shape):
        if self._C == 1:
            if len(shape) != 1:
                raise ValueError("single-channel blocks take 1-D arrays")
            return shape[0]
        want = "(n, C)" if self._layout == L.INTERLEAVED else "(C, n)"
        if len(shape) != 2 or shape[1 if self._layout == L.INTERLEAVED else 0] != self._C:
            raise ValueError(f"expected shape {want} with C = {self._C}, got {tuple(shape)}")
        return shape[0 if self._layout == L.INTERLEAVED else 1]

    def _out_shape(self, n_out, rate_changing):
        if self._C == 1:
            return (n_out,)
        if self._layout == L.INTERLEAVED and not rate_changing:
            return (n_out, self._C)
        return (self._C, n_out)  # CIC outputs are planar (b200dsp.h: b2d_cic_run)

    def _run(self, x, fn_host, fn_dev, max_out, rate_changing, out=None):
        lib = L.load()
        n_out = C.c_size_t(0)
        if _is_torch(x):
            import torch
            if not x.is_cuda:
                raise ValueError("torch inputs must live on the GPU (use numpy arrays for the host path)")
            tdt = {np.int16: torch.int16, np.int32: torch.int32, np.int64: torch.int64}
            if x.dtype != tdt[self._in_dt]:
                raise TypeError(f"input dtype must be {tdt[self._in_dt]}")
            x = x.contiguous()
            n = self._n_per_channel(tuple(x.shape))
            cap = max_out(n)
            if out is not None:
                if not out.is_cuda or out.dtype != tdt[self._out_dt] or out.numel() < self._C * cap or not out.is_contiguous():
                    raise ValueError("out= must be a contiguous CUDA tensor of the output container dtype and capacity")
                y = out.reshape(-1)[: self._C * cap].reshape(self._out_shape(cap, rate_changing))
            else:
                y = torch.empty(self._out_shape(cap, rate_changing), dtype=tdt[self._out_dt], device=x.device)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            L.check(fn_dev(self._h, x.data_ptr(), n, y.data_ptr(), C.byref(n_out), stream))
            if n_out.value != cap:  # planar stride is the true output count
                y = y.reshape(-1)[: self._C * n_out.value].reshape(self._out_shape(n_out.value, rate_changing))
            return y
        x = np.ascontiguousarray(np.asarray(x).astype(self._in_dt, copy=False))
        n = self._n_per_channel(x.shape)
        cap = max_out(n)
        if self._wire == L.WIRE_PACKED:
            pb = lib.b2d_wire_bytes(self._out_W, L.WIRE_PACKED)
            yb = np.empty(self._C * max(cap, 1) * pb, dtype=np.uint8)
            L.check(fn_host(self._h, x.ctypes.data, n, yb.ctypes.data, C.byref(n_out)))
            return yb[: self._C * n_out.value * pb].reshape(self._out_shape(n_out.value, rate_changing) + (pb,)).copy()
        y = np.empty(self._C * max(cap, 1), dtype=self._out_dt)
        L.check(fn_host(self._h, x.ctypes.data, n, y.ctypes.data, C.byref(n_out)))
        return y[: self._C * n_out.value].reshape(self._out_shape(n_out.value, rate_changing)).copy().view(self._host_dt)


class _Fir(_Block):
    _kind = "load"
    _prefix = "fir"

    def __init__(self, IN_TYPE, OUT_TYPE, COEFF_TYPE, ACC_TYPE, N_TAPS, ftype="SHIFT_REG", n_channels=1,
                 layout="planar", device=-1, comm=None, root=0):
        lib = L.load()
        self._h = None
        self.N_TAPS = int(N_TAPS)
        ft = L.FTYPES.index(ftype) if isinstance(ftype, str) else int(ftype)
        d = L.B2dFirDesc(L.make_fmt(IN_TYPE), L.make_fmt(COEFF_TYPE), L.make_fmt(ACC_TYPE), L.make_fmt(OUT_TYPE),
                         self.N_TAPS, ft, L.FIR_KINDS.index(self._kind), int(n_channels),
                         L.INTERLEAVED if layout in ("interleaved", L.INTERLEAVED) else L.PLANAR, int(device))
        h = C.c_void_p()
        L.check(lib.b2d_fir_create(C.byref(h), C.byref(d)))
        self._h = h
        self._coeff_dt = _container_dtype(d.coeff)
        self._setup_io(d.fin, d.out, n_channels, layout)
        if comm is not None:
            L.check(lib.b2d_fir_set_comm(self._h, comm._c, int(root)))

    @property
    def path(self):
        """Kernel family serving this configuration (b200dsp.h: b2d_fir_path)."""
        return L.load().b2d_fir_path(self._h).decode()

    def ovs_margin(self):
        """(a-priori error bound of the overlap-save evaluation for the loaded taps, largest distance from an integer seen so
        far or -1 when B2D_OVS_RESID=1 was not set at construction) -- b2d_fir_ovs_margin."""
        b, r = C.c_double(0), C.c_double(0)
        L.check(L.load().b2d_fir_ovs_margin(self._h, C.byref(b), C.byref(r)))
        return b.value, r.value

    def _load(self, coeffs, channel=-1):
        lib = L.load()
        if coeffs is None:  # non-root rank of a broadcast
            L.check(lib.b2d_fir_load(self._h, None, self.N_TAPS, int(channel)))
            return
        c = np.ascontiguousarray(np.asarray(coeffs).astype(self._coeff_dt, copy=False))
        L.check(lib.b2d_fir_load(self._h, c.ctypes.data, c.size, int(channel)))

    def _process(self, data_in, out=None):
        lib = L.load()
        return self._run(data_in, lib.b2d_fir_run, lib.b2d_fir_run_dev, lambda n: n, False, out)

    def reset(self):
        L.check(L.load().b2d_fir_reset(self._h))

    def get_state(self):
        lib = L.load()
        n = C.c_size_t(0)
        L.check(lib.b2d_fir_state_bytes(self._h, C.byref(n)))
        buf = np.zeros(n.value, dtype=np.uint8)
        L.check(lib.b2d_fir_get_state(self._h, buf.ctypes.data, buf.size))
        return buf

    def set_state(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        L.check(L.load().b2d_fir_set_state(self._h, blob.ctypes.data, blob.size))

    def close(self):
        if getattr(self, "_h", None):
            L.load().b2d_fir_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ac_fir_const_coeffs(_Fir):
    """ac_fir_const_coeffs<IN,OUT,COEFF,ACC,N_TAPS,ftype>(const COEFF *coeffs); run(in, out)
    (reference ac_fir_const_coeffs.h:309-355).  The coefficient array is bound at construction."""
    _kind = "const"

    def __init__(self, IN_TYPE, OUT_TYPE, COEFF_TYPE, ACC_TYPE, N_TAPS, ftype, coeffs, **kw):
        super().__init__(IN_TYPE, OUT_TYPE, COEFF_TYPE, ACC_TYPE, N_TAPS, ftype, **kw)
        self._load(coeffs)

    def run(self, data_in, out=None):
        return self._process(data_in, out)


class ac_fir_load_coeffs(_Fir):
    """ac_fir_load_coeffs<...>::run(data_in, coeffs_ch, data_out, ld)  (reference ac_fir_load_coeffs.h:300-365).
    One `ld` token is consumed per call; coefficients are taken only if ld is true AND N_TAPS values are queued,
    otherwise the token is silently dropped (:324-331)."""
    _kind = "load"

    def run(self, data_in=None, coeffs_ch=None, ld=None, channel=-1, out=None):
        if ld is not None and bool(ld) and coeffs_ch is not None and len(coeffs_ch) >= self.N_TAPS:
            self._load(np.asarray(coeffs_ch)[: self.N_TAPS], channel)
        if data_in is None or len(data_in) == 0:
            return np.empty(0, dtype=self._host_dt)
        return self._process(data_in, out)

    def load(self, coeffs, channel=-1):
        """Convenience: run() with ld=true and N_TAPS coefficients queued (multi-GPU: `coeffs` may be None off-root)."""
        self._load(coeffs, channel)


class ac_fir_prog_coeffs(_Fir):
    """ac_fir_prog_coeffs<...>::run(data_in, data_out, coeffs[N_TAPS])  (reference ac_fir_prog_coeffs.h:261-303).
    The reference consumes one sample per call; here a call consumes the whole array, which is the same as calling
    the reference once per sample with the same coefficient array.  The delay line survives a coefficient change."""
    _kind = "prog"

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._last = None

    def run(self, data_in, coeffs=None, channel=-1, out=None):
        if coeffs is not None:
            key = np.asarray(coeffs).tobytes()
            if key != self._last or channel >= 0:
                self._load(coeffs, channel)
                self._last = key if channel < 0 else None
        return self._process(data_in, out)

    def load(self, coeffs, channel=-1):
        self._load(coeffs, channel)
        self._last = None


class ac_fir_reg_share(_Fir):
    """ac_fir_reg_share<N_TAPS, IN, OUT, COEFF, ACC, MEM_WORD_WIDTH, BLK_SZ, BLK_OFFSET, ftype>
    (reference ac_fir_reg_share.h:257-303): SHIFT_REG / FOLD_EVEN / FOLD_ODD and the two
    anti-symmetric folds, taps read from a blocked coefficient RAM.  The reference's run() is scalar on a caller-owned
    delay line; run(samples, coeffs_ram) here is that call repeated over the array with the object's own (initially
    zero) delay line, delay_line() is ac_firProgCoeffs_delay_line, run_window() the scalar form on an explicit line."""
    _kind = "reg_share"

    def __init__(self, N_TAPS, IN_TYPE, OUT_TYPE, COEFF_TYPE, ACC_TYPE, MEM_WORD_WIDTH=1, BLK_SZ=1, BLK_OFFSET=0,
                 ftype="SHIFT_REG", **kw):
        super().__init__(IN_TYPE, OUT_TYPE, COEFF_TYPE, ACC_TYPE, N_TAPS, ftype, **kw)
        self._blk = (int(MEM_WORD_WIDTH), int(BLK_SZ), int(BLK_OFFSET))
        self._last = None

    def load(self, coeffs_ram, channel=-1):
        c = np.ascontiguousarray(np.asarray(coeffs_ram).astype(self._coeff_dt, copy=False))
        L.check(L.load().b2d_fir_load_blocked(self._h, c.ctypes.data, c.size, *self._blk, int(channel)))
        self._last = c.tobytes() if channel < 0 else None

    def run(self, data_in, coeffs_ram=None, out=None):
        if coeffs_ram is not None:
            c = np.ascontiguousarray(np.asarray(coeffs_ram).astype(self._coeff_dt, copy=False))
            if c.tobytes() != self._last:
                self.load(c)
        return self._process(data_in, out)

    def delay_line(self):
        y = np.zeros(self._C, dtype=self._out_dt)
        L.check(L.load().b2d_fir_delay_line_out(self._h, y.ctypes.data))
        y = y.view(self._host_dt)
        return y if self._C > 1 else y[0]

    def run_window(self, reg):
        w = np.ascontiguousarray(np.asarray(reg).astype(self._in_dt, copy=False))
        if w.size != self._C * self.N_TAPS:
            raise ValueError("window must hold n_channels * N_TAPS samples (reg[0] newest)")
        y = np.zeros(self._C, dtype=self._out_dt)
        L.check(L.load().b2d_fir_run_window(self._h, w.ctypes.data, y.ctypes.data))
        y = y.view(self._host_dt)
        return y if self._C > 1 else y[0]


class _Cic(_Block):
    _mode = 0

    _prefix = "cic"

    def __init__(self, IN_TYPE, OUT_TYPE, R, M, N, n_channels=1, layout="planar", device=-1):
        lib = L.load()
        self._h = None
        d = L.B2dCicDesc(L.make_fmt(IN_TYPE), L.make_fmt(OUT_TYPE), int(R), int(M), int(N), self._mode, int(n_channels),
                         L.INTERLEAVED if layout in ("interleaved", L.INTERLEAVED) else L.PLANAR, int(device))
        w = C.c_int32(0)
        L.check(lib.b2d_cic_int_width(C.byref(d), C.byref(w)))
        self.int_width = w.value
        h = C.c_void_p()
        L.check(lib.b2d_cic_create(C.byref(h), C.byref(d)))
        self._h = h
        self.R, self.M, self.N = int(R), int(M), int(N)
        self._setup_io(d.fin, d.out, n_channels, layout)

    @property
    def path(self):
        return L.load().b2d_cic_path(self._h).decode()

    def run(self, data_in, out=None):
        lib = L.load()
        return self._run(data_in, lib.b2d_cic_run, lib.b2d_cic_run_dev, lambda n: lib.b2d_cic_max_out(self._h, n), True, out)

    def reset(self):
        L.check(L.load().b2d_cic_reset(self._h))

    def get_state(self):
        lib = L.load()
        n = C.c_size_t(0)
        L.check(lib.b2d_cic_state_bytes(self._h, C.byref(n)))
        buf = np.zeros(n.value, dtype=np.uint8)
        L.check(lib.b2d_cic_get_state(self._h, buf.ctypes.data, buf.size))
        return buf

    def set_state(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        L.check(L.load().b2d_cic_set_state(self._h, blob.ctypes.data, blob.size))

    def close(self):
        if getattr(self, "_h", None):
            L.load().b2d_cic_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ac_cic_dec_full(_Cic):
    """ac_cic_dec_full<IN, OUT, R, M, N>::run(data_in, data_out)  (reference ac_cic_dec_full.h:147-222)."""
    _mode = 0


class ac_cic_intr_full(_Cic):
    """ac_cic_intr_full<IN, OUT, R, M, N>::run(data_in, data_out)  (reference ac_cic_intr_full.h:137-215)."""
    _mode = 1


class cic_intr_fir_cascade(_Block):
    """ac_cic_intr_full<IN, MID, R, M, N> feeding ac_fir_*<MID, OUT, COEFF, ACC, N_TAPS, ftype> as ONE object
    (BASELINE config 5): run(data_in) == fir.run(cic.run(data_in)) of the two reference objects, bit for bit, with the
    same stream-edge behaviour.  When no stage drops bits the engine fuses both into a single polyphase kernel on the
    16-bit input (path 'cicfir_fused'); otherwise the two kernels run back to back on the device."""

    _prefix = "cicfir"

    def __init__(self, IN_TYPE, MID_TYPE, R, M, N, OUT_TYPE, COEFF_TYPE, ACC_TYPE, N_TAPS, ftype="SHIFT_REG", coeffs=None,
                 n_channels=1, layout="planar", device=-1, fir_class="load"):
        lib = L.load()
        self._h = None
        self.N_TAPS = int(N_TAPS)
        lay = L.INTERLEAVED if layout in ("interleaved", L.INTERLEAVED) else L.PLANAR
        ft = L.FTYPES.index(ftype) if isinstance(ftype, str) else int(ftype)
        cd = L.B2dCicDesc(L.make_fmt(IN_TYPE), L.make_fmt(MID_TYPE), int(R), int(M), int(N), 1, int(n_channels), lay, int(device))
        fd = L.B2dFirDesc(L.make_fmt(MID_TYPE), L.make_fmt(COEFF_TYPE), L.make_fmt(ACC_TYPE), L.make_fmt(OUT_TYPE),
                          self.N_TAPS, ft, L.FIR_KINDS.index(fir_class), int(n_channels), L.PLANAR, int(device))
        h = C.c_void_p()
        L.check(lib.b2d_cicfir_create(C.byref(h), C.byref(cd), C.byref(fd)))
        self._h = h
        self._coeff_dt = _container_dtype(fd.coeff)
        self._setup_io(cd.fin, fd.out, n_channels, layout)
        if coeffs is not None:
            self.load(coeffs)

    @property
    def path(self):
        return L.load().b2d_cicfir_path(self._h).decode()

    def load(self, coeffs, channel=-1):
        c = np.ascontiguousarray(np.asarray(coeffs).astype(self._coeff_dt, copy=False))
        L.check(L.load().b2d_cicfir_load(self._h, c.ctypes.data, c.size, int(channel)))

    def run(self, data_in, out=None):
        lib = L.load()
        return self._run(data_in, lib.b2d_cicfir_run, lib.b2d_cicfir_run_dev, lambda n: lib.b2d_cicfir_max_out(self._h, n), True, out)

    def get_state(self):
        """Checkpoint blob (b2d_cicfir_get_state)."""
        return _get_state(self._h, "cicfir")

    def set_state(self, blob):
        _set_state(self._h, "cicfir", blob)

    def reset(self):
        L.check(L.load().b2d_cicfir_reset(self._h))

    def close(self):
        if getattr(self, "_h", None):
            L.load().b2d_cicfir_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ac_poly_dec(_Block):
    """ac_poly_dec<IN, COEFF, STR_COEFF, ACC, OUT, NTAPS, DF>::run(data_in, data_out, coeffs_st)
    (reference ac_poly_dec.h:87-137): polyphase decimator, one output per DF inputs, coefficients
    coeffs[NTAPS * DF] in phase order.  A call consumes whole groups of DF samples; the rest stays pending."""

    _prefix = "polydec"

    def __init__(self, IN_TYPE, COEFF_TYPE, ACC_TYPE, OUT_TYPE, NTAPS, DF, coeffs=None, n_channels=1, layout="planar", device=-1):
        lib = L.load()
        self._h = None
        self.NTAPS, self.DF = int(NTAPS), int(DF)
        d = L.B2dPolydecDesc(L.make_fmt(IN_TYPE), L.make_fmt(COEFF_TYPE), L.make_fmt(ACC_TYPE), L.make_fmt(OUT_TYPE), self.NTAPS, self.DF,
                             int(n_channels), L.INTERLEAVED if layout in ("interleaved", L.INTERLEAVED) else L.PLANAR, int(device))
        h = C.c_void_p()
        L.check(lib.b2d_polydec_create(C.byref(h), C.byref(d)))
        self._h = h
        self._coeff_dt = _container_dtype(d.coeff)
        self._setup_io(d.fin, d.out, n_channels, layout)
        if coeffs is not None:
            self.load(coeffs)

    @property
    def path(self):
        return L.load().b2d_polydec_path(self._h).decode()

    def load(self, coeffs, channel=-1):
        c = np.ascontiguousarray(np.asarray(coeffs).astype(self._coeff_dt, copy=False))
        L.check(L.load().b2d_polydec_load(self._h, c.ctypes.data, c.size, int(channel)))

    def run(self, data_in, coeffs_st=None, out=None):
        if coeffs_st is not None:
            self.load(coeffs_st)
        lib = L.load()
        return self._run(data_in, lib.b2d_polydec_run, lib.b2d_polydec_run_dev, lambda n: lib.b2d_polydec_max_out(self._h, n), True, out)

    def get_state(self):
        """Checkpoint blob (b2d_polydec_get_state)."""
        return _get_state(self._h, "polydec")

    def set_state(self, blob):
        _set_state(self._h, "polydec", blob)

    def reset(self):
        L.check(L.load().b2d_polydec_reset(self._h))

    def close(self):
        if getattr(self, "_h", None):
            L.load().b2d_polydec_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ac_poly_intr(_Block):
    """ac_poly_intr<IN, COEFF, ACC, OUT, STR_CTRL, STR_COEFF, NTAPS, COEFFSZ, IF, ftype>::run(data_in, data_out, ctrl_st,
    coeffs_st, read_ctrl_chan)  (reference ac_poly_intr.h:261-312): polyphase interpolator, IF
    outputs per input.  ftype is "FOLD_EVEN" / "FOLD_ODD" (symmetric-pair structures: outputs one step late, sign[] and
    corr[] from the control struct) or "FOLD_ANTI" (plain polyphase form).  load() is the read_ctrl = true call."""

    _prefix = "polyintr"

    def __init__(self, IN_TYPE, COEFF_TYPE, ACC_TYPE, OUT_TYPE, NTAPS, IF, ftype="FOLD_ANTI", coeffs=None, sign=None, corr=None,
                 n_channels=1, layout="planar", device=-1):
        lib = L.load()
        self._h = None
        self.NTAPS, self.IF = int(NTAPS), int(IF)
        ft = L.PI_FTYPES.index(ftype) if isinstance(ftype, str) else int(ftype)
        d = L.B2dPolyintrDesc(L.make_fmt(IN_TYPE), L.make_fmt(COEFF_TYPE), L.make_fmt(ACC_TYPE), L.make_fmt(OUT_TYPE), self.NTAPS, self.IF, ft,
                              int(n_channels), L.INTERLEAVED if layout in ("interleaved", L.INTERLEAVED) else L.PLANAR, int(device))
        h = C.c_void_p()
        L.check(lib.b2d_polyintr_create(C.byref(h), C.byref(d)))
        self._h = h
        self._coeff_dt = _container_dtype(d.coeff)
        self._setup_io(d.fin, d.out, n_channels, layout)
        if coeffs is not None:
            self.load(coeffs, sign, corr)

    @property
    def path(self):
        return L.load().b2d_polyintr_path(self._h).decode()

    @property
    def coeffsz(self):
        return int(L.load().b2d_polyintr_coeffsz(self._h))

    def load(self, coeffs, sign=None, corr=None, channel=-1):
        c = np.ascontiguousarray(np.asarray(coeffs).astype(self._coeff_dt, copy=False))
        sg = None if sign is None else np.ascontiguousarray(np.asarray(sign) != 0, dtype=np.uint8)
        cr = None if corr is None else np.ascontiguousarray(np.asarray(corr), dtype=np.uint8)
        for a in (sg, cr):
            if a is not None and a.size != self.IF:
                raise ValueError("sign / corr need IF entries")
        L.check(L.load().b2d_polyintr_load(self._h, c.ctypes.data, c.size, None if sg is None else sg.ctypes.data,
                                           None if cr is None else cr.ctypes.data, int(channel)))

    def run(self, data_in, out=None):
        lib = L.load()
        return self._run(data_in, lib.b2d_polyintr_run, lib.b2d_polyintr_run_dev, lambda n: lib.b2d_polyintr_max_out(self._h, n), True, out)

    def get_state(self):
        """Checkpoint blob (b2d_polyintr_get_state)."""
        return _get_state(self._h, "polyintr")

    def set_state(self, blob):
        _set_state(self._h, "polyintr", blob)

    def reset(self):
        L.check(L.load().b2d_polyintr_reset(self._h))

    def close(self):
        if getattr(self, "_h", None):
            L.load().b2d_polyintr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ac_mv_avg:
    """ac_mv_avg<MAX_SAMPLE, TAPS, WIN_TYPE, IN, OUT, ACC, COEFF, S_TYPE>(coeffs)::run(data_in, data_out, n_sample)
    (reference ac_mv_avg.h:140-204): run(whole bursts of n_sample samples, n_sample) -> the smoothed
    bursts (n_sample outputs each for AC_CLIP / AC_MIRROR, n_sample - TAPS + 1 for AC_WIN).  Parity unpinned: the window
    class lives in ac_math (absent); its boundary rules are restated from the manual."""

    def __init__(self, MAX_SAMPLE, TAPS, WIN_TYPE, IN_TYPE, OUT_TYPE, ACC_TYPE, COEFF_TYPE, coeffs, device=-1):
        lib = L.load()
        self._h = None
        w = L.WIN_MODES.index(WIN_TYPE) if isinstance(WIN_TYPE, str) else int(WIN_TYPE)
        d = L.B2dMvavgDesc(L.make_fmt(IN_TYPE), L.make_fmt(OUT_TYPE), L.make_fmt(ACC_TYPE), L.make_fmt(COEFF_TYPE), int(MAX_SAMPLE), int(TAPS),
                           w, int(device))
        c = np.ascontiguousarray(np.asarray(coeffs).astype(_container_dtype(d.coeff), copy=False))
        if c.size != int(TAPS):
            raise ValueError("expected TAPS coefficients")
        h = C.c_void_p()
        L.check(lib.b2d_mvavg_create(C.byref(h), C.byref(d), c.ctypes.data))
        self._h = h
        self._in_dt, self._out_dt, self._host_dt = _container_dtype(d.fin), _container_dtype(d.out), _host_dtype(d.out)

    def run(self, data_in, n_sample):
        lib = L.load()
        n_out = C.c_size_t(0)
        if _is_torch(data_in):
            import torch
            x = data_in.contiguous()
            y = torch.empty(max(x.numel(), 1), dtype={np.int16: torch.int16, np.int32: torch.int32, np.int64: torch.int64}[self._out_dt], device=x.device)
            L.check(lib.b2d_mvavg_run_dev(self._h, x.data_ptr(), x.numel(), int(n_sample), y.data_ptr(), C.byref(n_out),
                                          torch.cuda.current_stream(x.device).cuda_stream))
            return y[: n_out.value]
        x = np.ascontiguousarray(np.asarray(data_in).astype(self._in_dt, copy=False)).reshape(-1)
        y = np.empty(max(x.size, 1), dtype=self._out_dt)
        L.check(lib.b2d_mvavg_run(self._h, x.ctypes.data, x.size, int(n_sample), y.ctypes.data, C.byref(n_out)))
        return y[: n_out.value].copy().view(self._host_dt)

    @property
    def path(self):
        return L.load().b2d_mvavg_path(self._h).decode()

    def close(self):
        if getattr(self, "_h", None):
            L.load().b2d_mvavg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ac_intg_dump:
    """ac_intg_dump<IN, ACC, OUT, N_TYPE, NS, CHN>::run(data_in, data_out, n_sample)  (reference ac_intg_dump.h:113-151,
    SURVEY.md 8f row N4): run(samples interleaved over CHN, n_sample tokens) -> (dumping frames, CHN) sums."""

    def __init__(self, IN_TYPE, ACC_TYPE, OUT_TYPE, NS, CHN, device=-1):
        lib = L.load()
        self._h = None
        d = L.B2dIntgdumpDesc(L.make_fmt(IN_TYPE), L.make_fmt(ACC_TYPE), L.make_fmt(OUT_TYPE), int(NS), int(CHN), int(device))
        h = C.c_void_p()
        L.check(lib.b2d_intgdump_create(C.byref(h), C.byref(d)))
        self._h = h
        self.NS, self.CHN = int(NS), int(CHN)
        self._in_dt, self._out_dt, self._host_dt = _container_dtype(d.fin), _container_dtype(d.out), _host_dtype(d.out)

    def run(self, data_in, n_sample):
        lib = L.load()
        ns = np.ascontiguousarray(np.asarray(n_sample), dtype=np.uint32)
        n_out = C.c_size_t(0)
        if _is_torch(data_in):
            import torch
            x = data_in.contiguous()
            y = torch.empty((ns.size, self.CHN), dtype={np.int16: torch.int16, np.int32: torch.int32, np.int64: torch.int64}[self._out_dt], device=x.device)
            L.check(lib.b2d_intgdump_run_dev(self._h, x.data_ptr(), x.numel(), ns.ctypes.data, ns.size, y.data_ptr(), C.byref(n_out),
                                             torch.cuda.current_stream(x.device).cuda_stream))
            return y[: n_out.value // self.CHN]
        x = np.ascontiguousarray(np.asarray(data_in).astype(self._in_dt, copy=False)).reshape(-1)
        y = np.empty((max(ns.size, 1), self.CHN), dtype=self._out_dt)
        L.check(lib.b2d_intgdump_run(self._h, x.ctypes.data, x.size, ns.ctypes.data, ns.size, y.ctypes.data, C.byref(n_out)))
        return y[: n_out.value // self.CHN].copy().view(self._host_dt)

    @property
    def path(self):
        return L.load().b2d_intgdump_path(self._h).decode()

    def get_state(self):
        """Checkpoint blob (b2d_intgdump_get_state)."""
        return _get_state(self._h, "intgdump")

    def set_state(self, blob):
        _set_state(self._h, "intgdump", blob)

    def reset(self):
        L.check(L.load().b2d_intgdump_reset(self._h))

    def close(self):
        if getattr(self, "_h", None):
            L.load().b2d_intgdump_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """NCCL communicator of the C-ABI (one rank per GPU); only used for the coefficient broadcast at load()."""

    def __init__(self, unique_id, rank, world, device=-1):
        lib = L.load()
        self._c = C.c_void_p()
        idb = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        L.check(lib.b2d_comm_create(C.byref(self._c), idb, int(rank), int(world), int(device)))
        self.rank, self.world = int(rank), int(world)

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        L.check(L.load().b2d_comm_unique_id(buf))
        return bytes(buf)

    def barrier(self):
        L.check(L.load().b2d_comm_barrier(self._c))

    def close(self):
        if getattr(self, "_c", None):
            L.load().b2d_comm_destroy(self._c)
            self._c = None


def shard_channels(n_channels, rank, world):
    """Channel c lives on rank c % world (b2d_shard_count): the local channel ids of `rank`."""
    n = C.c_uint32(0)
    L.check(L.load().b2d_shard_count(int(n_channels), int(rank), int(world), C.byref(n)))
    ids = list(range(rank, n_channels, world))
    assert len(ids) == n.value
    return ids
