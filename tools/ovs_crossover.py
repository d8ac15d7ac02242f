#!/usr/bin/env python
"""tools/ovs_crossover.py -- where a call becomes long enough for the overlap-save evaluation: device-resident time of one
b2d_fir_run_dev call (CUDA events, 50 repetitions) through the DP2A kernel (B2D_FIR_OVS=0) and through fir_ovs
(B2D_FIR_OVS=2) over a range of call lengths; the engine's own choice (rt_fir.cu: fir_ovs_worth) is printed beside them."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(mode, taps, C, layout, n):
    if mode is None:
        os.environ.pop("B2D_FIR_OVS", None)
    else:
        os.environ["B2D_FIR_OVS"] = mode
    import ac_dsp_b200 as E
    rng = np.random.default_rng(1)
    h = rng.integers(-32768, 32767, size=taps, endpoint=True).astype(np.int16)
    f = E.ac_fir_load_coeffs((16, 1), (40, 8), (16, 1), (40, 8), taps, "SHIFT_REG", n_channels=C, layout=layout)
    f.load(h)
    shape = (n, 2) if layout == "interleaved" else (C, n)
    x = torch.randint(-32768, 32768, shape, dtype=torch.int16, device="cuda")
    y = torch.empty(shape, dtype=torch.int64, device="cuda")
    for _ in range(5):
        f.run(x, out=y.reshape(-1))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        f.run(x, out=y.reshape(-1))
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / 50 * 1e3


for taps, C, layout in ((256, 2, "interleaved"), (1024, 8, "planar"), (128, 2, "interleaved")):
    for n in (20000, 50000, 100000, 200000, 300000, 500000, 1000000, 4000000):
        if layout == "planar":
            n //= 4
        t0, t2, ta = timed("0", taps, C, layout, n), timed("2", taps, C, layout, n), timed(None, taps, C, layout, n)
        print(json.dumps({"taps": taps, "channels": C, "layout": layout, "n_per_channel": n, "us_dp2a": round(t0, 1), "us_ovs": round(t2, 1),
                          "us_engine_choice": round(ta, 1), "engine_picks": "ovs" if abs(ta - t2) < abs(ta - t0) else "dp2a"}))
