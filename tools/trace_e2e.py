import sys, os, ctypes as ct, numpy as np, time
sys.path.insert(0, '/root/repo' if os.path.exists('/root/repo/bench.py') else '.')
import torch
import ac_dsp_b200 as E
from bench import host_buffer
lib = E.load()
n2 = 1 << 27
h = np.random.default_rng(1).integers(-32768, 32767, 256).astype(np.int16)
f = E.ac_fir_load_coeffs((16,1),(40,8),(16,1),(40,8),256,"SHIFT_REG",n_channels=2,layout="interleaved")
f.load(h)
xb, xp = host_buffer(lib, n2*4); yb, yp = host_buffer(lib, n2*16)
xb[:] = 7
no = ct.c_size_t(0)
for wire in (0, 1):
    lib.b2d_fir_set_wire(f._h, wire)
    for rep in range(3):
        if rep == 2: os.environ["B2D_PIPE_TRACE"] = "1"
        t0 = time.perf_counter()
        assert lib.b2d_fir_run(f._h, xb.ctypes.data, n2, yb.ctypes.data, ct.byref(no)) == 0
        dt = time.perf_counter() - t0
        os.environ.pop("B2D_PIPE_TRACE", None)
        print("wire", wire, "rep", rep, "ms %.2f" % (dt*1e3), "M IQ/s %.0f" % (n2/dt/1e6), file=sys.stderr)
