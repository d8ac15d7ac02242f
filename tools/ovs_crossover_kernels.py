#!/usr/bin/env python
"""tools/ovs_crossover_kernels.py -- one call per (kernel family, call length), to be run under
`ncu --metrics gpu__time_duration.sum`: the launch list gives the GPU-side duration of fir_q15_kernel and fir_ovs_kernel
without the host-side call overhead that hides it in a Python timing loop (tools/ovs_crossover.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for taps, C, layout, sizes in ((256, 2, "interleaved", (100000, 300000, 600000, 1200000, 2400000)),
                               (1024, 8, "planar", (12500, 25000, 50000, 100000, 250000))):
    for n in sizes:
        for mode in ("0", "2"):
            os.environ["B2D_FIR_OVS"] = mode
            import ac_dsp_b200 as E
            rng = np.random.default_rng(1)
            h = rng.integers(-32768, 32767, size=taps, endpoint=True).astype(np.int16)
            f = E.ac_fir_load_coeffs((16, 1), (40, 8), (16, 1), (40, 8), taps, "SHIFT_REG", n_channels=C, layout=layout)
            f.load(h)
            shape = (n, 2) if layout == "interleaved" else (C, n)
            x = torch.randint(-32768, 32768, shape, dtype=torch.int16, device="cuda")
            for _ in range(3):
                y = f.run(x)
            torch.cuda.synchronize()
            print(taps, C, layout, n, mode, flush=True)
