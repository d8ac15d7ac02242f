#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on the B200 engine, with the reference's CPU path beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload fir256|fir1024|cic_dec|cic_intr] [--impl reference]

A step is one run() of the hot path over one batch of synthetic 16-bit samples already resident in HBM
(default workload: BASELINE.json configs[1], the 256-tap ac_fixed<16,1> -> <40,8> FIR over 2^30 interleaved IQ
samples per GPU).  Rank 0 prints ONE JSON line.  `value` is device-resident throughput (CUDA events, max over
ranks); `e2e` is the same metric through the C-ABI host-buffer call (b2d_*_run on pinned host memory, copies
inside the timed region); `roofline` is the dominant kernel's algorithmic HBM bytes / its event-timed duration
against MEASURED_PEAKS.json; `cpu_baseline` is the reference's own C++ templates (oracle/_ref, built in the dev
container from /root/reference over the clean-room ac_types shim) on this box's host cores over a bounded sample.
`--impl reference` prints that CPU run as its own line.  The oracle is only ever the thing timed as the
CPU baseline here -- never part of the GPU path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line: libraries that print banners on fd 1 (NCCL's version line) are sent to stderr
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())

Q15, ACC40 = (16, 1), (40, 8)
SEED = 20260101

# name -> description of one step per GPU.  bytes_per_unit: SURVEY.md 8(d) algorithmic bytes per counted sample.
WORKLOADS = {
    # BASELINE.json configs[1]
    "fir256": dict(kind="fir", taps=256, channels=2, layout="interleaved", n=1 << 30, unit_is_iq=True,
                   bytes_per_unit=20.0, macs_per_unit=512,
                   name="ac_fir_load_coeffs 256-tap ac_fixed<16,1,true> x <16,1,true> -> <40,8,true>, interleaved 16-bit IQ, 2^30 IQ samples per GPU"),
    # BASELINE.json configs[3], per-GPU share: 8 real channels x 2^27 samples, 1024 taps
    "fir1024": dict(kind="fir", taps=1024, channels=8, layout="planar", n=1 << 27, unit_is_iq=False,
                    bytes_per_unit=10.0, macs_per_unit=1024,
                    name="ac_fir_prog_coeffs 1024-tap <16,1> -> <40,8>, 8 real channels x 2^27 samples per GPU"),
    # BASELINE.json configs[2]
    "cic_dec": dict(kind="cic", mode="dec", R=8, M=1, N=4, out=(28, 13), channels=2, layout="interleaved", n=1 << 30,
                    unit_is_iq=True, bytes_per_unit=5.0, macs_per_unit=0,
                    name="ac_cic_dec_full R=8 M=1 N=4 ac_fixed<16,1,true> -> <28,13,true>, interleaved 16-bit IQ, 2^30 IQ inputs per GPU"),
    # BASELINE.json configs[4] second stage, unfused: the wide (IMAD.WIDE) path on the interpolator's <20,5> output
    "fir63": dict(kind="fir", taps=63, channels=1, layout="planar", n=1 << 28, unit_is_iq=False, infmt=(20, 5),
                  bytes_per_unit=12.0, macs_per_unit=63,
                  name="ac_fir_const_coeffs 63-tap <20,5> x <16,1> -> <40,8>, 1 real channel x 2^28 samples per GPU (int32 in, int64 out)"),
    # BASELINE.json configs[4], per-GPU share (1 real channel): interpolator + 63-tap FIR as one fused polyphase kernel
    "cicfir": dict(kind="cicfir", R=4, M=1, N=3, mid=(20, 5), taps=63, channels=1, layout="planar", n=1 << 26,
                   unit_is_iq=False, bytes_per_unit=34.0, macs_per_unit=4 * 18 * 1.5,
                   name="ac_cic_intr_full R=4 M=1 N=3 <16,1> -> <20,5> + 63-tap FIR <20,5> x <16,1> -> <40,8>, fused, 1 real channel x 2^26 inputs per GPU"),
    # the decimate-by-8 polyphase FIR that follows the R = 8 CIC decimator in a DDC chain
    "polydec": dict(kind="polydec", taps=32, df=8, channels=2, layout="interleaved", n=1 << 30, unit_is_iq=True,
                    bytes_per_unit=6.0, macs_per_unit=64,
                    name="ac_poly_dec NTAPS=32 DF=8 (256 taps) <16,1> x <16,1> -> <40,8>, interleaved 16-bit IQ, 2^30 IQ inputs per GPU"),
    # polyphase interpolator, plain form, 16 taps per phase x 4 phases (64-tap prototype)
    "polyintr": dict(kind="polyintr", taps=16, IF=4, channels=1, layout="planar", n=1 << 26, unit_is_iq=False,
                     bytes_per_unit=2.0 + 4 * 8.0, macs_per_unit=64,
                     name="ac_poly_intr NTAPS=16 IF=4 FOLD_ANTI (64-tap prototype) <16,1> x <16,1> -> <40,8>, 1 real channel x 2^26 inputs per GPU"),
    # integrate-and-dump, 4 interleaved channels, 64 samples per dump
    "intgdump": dict(kind="intgdump", chn=4, nsamp=256, ns=1024, channels=1, layout="planar", n=1 << 30, unit_is_iq=False,
                     bytes_per_unit=2.0 + 4.0 / 256, macs_per_unit=0,
                     name="ac_intg_dump CHN=4, 256 samples per dump, <16,1> -> <32,17>, 2^30 samples per GPU"),
    # BASELINE.json configs[4] first stage
    "cic_intr": dict(kind="cic", mode="intr", R=4, M=1, N=3, out=(20, 5), channels=1, layout="planar", n=1 << 28,
                     unit_is_iq=False, bytes_per_unit=18.0, macs_per_unit=0,
                     name="ac_cic_intr_full R=4 M=1 N=3 <16,1> -> <20,5>, 1 real channel x 2^28 inputs per GPU"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, dev):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        out, _ = self.p.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------- CPU reference
def cpu_reference(wl, seconds_target=12.0, threads=None):
    """The reference's own run() for this workload on the host cores: one filter instance per thread (an instance is
    inherently sequential; instances are independent, the same decomposition as the GPU's channel sharding)."""
    from oracle import oracle as O
    O.build()
    threads = threads or len(os.sched_getaffinity(0))
    kind = "reference" if O.have_ref() else "port"
    rng = np.random.default_rng(SEED)
    if wl["kind"] == "fir":
        taps = wl["taps"]
        h = O.rand_raw(rng, Q15, taps)

        def make():
            fi = wl.get("infmt", Q15)
            f = (O.FirA("load", fi, Q15, ACC40, ACC40, taps, "SHIFT_REG") if kind == "reference"
                 else O.FirB(fi, Q15, ACC40, ACC40, taps, "SHIFT_REG"))
            f.load(h)
            return f
        per_thread = int(seconds_target * 0.5e6 * 256 / taps)       # ~0.5 M real samples/s/core at 256 taps
    elif wl["kind"] == "intgdump":
        class IdRun:
            def __init__(self):
                self.f = O.IdA(0) if kind == "reference" else O.IdB(Q15, (32, 17), (32, 17), 64, wl["chn"])
            def run(self, x):
                # the compiled-in reference instantiation (ref_configs.ID_CONFIGS[0]) has NS = 64: the CPU arm dumps every
                # 64 samples (same adds per sample, 4x the dumps of the GPU workload)
                per = 64 if kind == "reference" else wl["nsamp"]
                m = (len(x) // (wl["chn"] * per)) * wl["chn"] * per
                return self.f.run(x[:m], np.full(m // (wl["chn"] * per), per))
            def last_run_seconds(self):
                return None
        make = IdRun
        per_thread = int(seconds_target * 8e6)
    elif wl["kind"] == "polyintr":
        cid = [i for i, c in enumerate(O.rc.PI_CONFIGS) if c[4] == wl["taps"] and c[5] == wl["IF"] and c[6] == "FOLD_ANTI" and c[0][0] == 16][0]
        h = O.rand_raw(rng, Q15, wl["taps"] * wl["IF"])

        def make():
            f = O.PiA(cid) if kind == "reference" else O.PiB(Q15, Q15, ACC40, ACC40, wl["taps"], wl["IF"], "FOLD_ANTI")
            f.load(h)
            f.last_run_seconds = lambda: None
            return f
        per_thread = int(seconds_target * 0.5e6)
    elif wl["kind"] == "polydec":
        cid = [i for i, c in enumerate(O.rc.PD_CONFIGS) if c[4] == wl["taps"] and c[5] == wl["df"] and c[0][0] == 16][0]
        h = O.rand_raw(rng, Q15, wl["taps"] * wl["df"])

        def make():
            f = O.PdA(cid) if kind == "reference" else O.PdB(Q15, Q15, ACC40, ACC40, wl["taps"], wl["df"])
            f.load(h)
            f.last_run_seconds = lambda: None
            return f
        per_thread = int(seconds_target * 2e6)
    elif wl["kind"] == "cicfir":
        taps = wl["taps"]
        h = O.rand_raw(rng, Q15, taps)

        class Chain:   # cic.run(in, mid); fir.run(mid, out) -- the two reference objects joined by a channel
            def __init__(self):
                self.cic = (O.CicA if kind == "reference" else O.CicB)("intr", Q15, wl["mid"], wl["R"], wl["M"], wl["N"])
                self.fir = (O.FirA("load", wl["mid"], Q15, ACC40, ACC40, taps, "SHIFT_REG") if kind == "reference"
                            else O.FirB(wl["mid"], Q15, ACC40, ACC40, taps, "SHIFT_REG"))
                self.fir.load(h)
                self.secs = 0.0

            def run(self, x):
                mid = self.cic.run(x)
                y = self.fir.run(mid)
                if kind == "reference":
                    self.secs = self.cic.last_run_seconds() + self.fir.last_run_seconds()
                return y

            def last_run_seconds(self):
                return self.secs
        make = Chain
        per_thread = int(seconds_target * 2e6 * 63 / taps / wl["R"])
    else:
        def make():
            cls = O.CicA if kind == "reference" else O.CicB
            return cls(wl["mode"], Q15, wl["out"], wl["R"], wl["M"], wl["N"])
        per_thread = min(1 << 23, int(seconds_target * (20e6 if wl["mode"] == "dec" else 4e6)))   # bounded: 16 B per queued sample
    per_thread = max(1 << 12, per_thread)
    threads = min(threads, 64)
    objs = [make() for _ in range(threads)]
    xs = [O.rand_raw(rng, wl.get("infmt", Q15), per_thread) for _ in range(threads)]
    for o in objs:
        o.run(xs[0][:2048])                                          # warm caches / page in

    secs = [0.0] * threads

    passes = [0] * threads

    def work(i):
        # the block is filtered again (the object's state simply continues) until the bounded sample has cost about
        # seconds_target of CPU time per thread; the hot loop got faster than the constants above assumed
        while passes[i] < 8 and secs[i] < 0.7 * seconds_target:
            t = time.perf_counter()
            objs[i].run(xs[i])
            # the reference arm times run() alone (channels pre-filled, drained afterwards): BASELINE.md section 3
            inner = objs[i].last_run_seconds() if kind == "reference" else None
            secs[i] += inner if inner else time.perf_counter() - t
            passes[i] += 1
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    rate_real = sum(per_thread * p / s for p, s in zip(passes, secs))   # instances run concurrently: rates add
    rate = rate_real / 2 if wl["unit_is_iq"] else rate_real
    dt = max(secs)
    return {"value": rate / 1e6, "unit": "Msamples/s", "cores": threads, "kind": kind, "seconds": dt,
            "sample": f"{threads} independent filter instances (one per host thread) x {per_thread} real samples x {max(passes)} pass(es) each, "
                      f"{'reference C++ templates over the ac_types shim (oracle/_ref)' if kind == 'reference' else 'oracle_b.c integer restatement'}"}


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_reference(wl, seconds_target=args.ref_seconds)
        if i >= args.warmup:
            vals.append(last)
    tot_s = sum(v["seconds"] for v in vals)
    v = float(np.mean([v["value"] for v in vals]))
    line = {"impl": "reference", "metric": "Msamples/s (16b IQ, 256-tap FIR)" if args.workload == "fir256" else f"Msamples/s ({args.workload})",
            "value": v, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_s / max(1, len(vals)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "ac_fixed (exact integer)", "data": "synthetic",
            "config": {"workload": wl["name"], "note": "CPU reference arm: bounded sample per step, host cores only"},
            "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": last["cores"], "kind": last["kind"], "sample": last["sample"]},
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def bind_near_gpu(local):
    """Run this rank's host threads -- and, by first touch, its pinned staging buffers -- on the CPUs next to its GPU
    (NVML's ideal affinity), so that the H2D / D2H streams of N ranks do not all cross the socket interconnect.
    Returns the original affinity (restored before the CPU baseline, which uses every core) or None."""
    try:
        import pynvml
        orig = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1} & orig
        if cpus:
            os.sched_setaffinity(0, cpus)
            return orig
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="fir256", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=None, help="override samples per channel per step (power of two)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ref-seconds", type=float, default=2.0, help="--impl reference: CPU seconds per step (bounded sample)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.log2n:
        wl["n"] = 1 << args.log2n
    if args.impl == "reference":
        return run_reference(args, wl)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    orig_affinity = bind_near_gpu(local) if world > 1 else None
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ac_dsp_b200 as E
    from ac_dsp_b200 import build as _b
    if rank == 0:
        _b.build()
    if world > 1:
        dist.barrier()
    E.load()

    # ---- coefficient set: rank 0 owns it, one ncclBroadcast at load() (the only collective on this path)
    from ac_dsp_b200 import parallel as P
    comm = P.make_comm(rank, world, local)
    rng = np.random.default_rng(SEED)
    C, n, il = wl["channels"], wl["n"], wl["layout"] == "interleaved"
    gen = torch.Generator(device="cuda").manual_seed(SEED + rank)
    shape = (n, C) if il else ((C, n) if C > 1 else (n,))
    infmt = wl.get("infmt", Q15)
    lim = 1 << (infmt[0] - 1)
    x = torch.randint(-lim, lim, shape, dtype=torch.int16 if infmt[0] <= 16 else torch.int32, device="cuda", generator=gen)
    if wl["kind"] == "fir":
        h = rng.integers(-32768, 32767, size=wl["taps"], endpoint=True).astype(np.int16)
        f = E.ac_fir_load_coeffs(infmt, ACC40, Q15, ACC40, wl["taps"], "SHIFT_REG", n_channels=C, layout=wl["layout"],
                                 device=local, comm=comm, root=0)
        f.load(h if rank == 0 else None)
        launches_per_step = 2          # fir_q15_kernel + history carry
    elif wl["kind"] == "intgdump":
        class _Id:   # adapter: fixed token array, out= ignored (outputs are 1/64 of the input)
            def __init__(self):
                self.f = E.ac_intg_dump(Q15, (32, 17), (32, 17), wl["ns"], wl["chn"], device=local)
                self._h = self.f._h
                self.tok = np.full(n // (wl["chn"] * wl["nsamp"]), wl["nsamp"], dtype=np.uint32)
            def run(self, x, out=None):
                return self.f.run(x, self.tok)
            @property
            def path(self):
                return self.f.path
            def close(self):
                self.f.close()
        f = _Id()
        launches_per_step = 1
    elif wl["kind"] == "polyintr":
        h = rng.integers(-32768, 32767, size=wl["taps"] * wl["IF"], endpoint=True).astype(np.int16)
        f = E.ac_poly_intr(Q15, Q15, ACC40, ACC40, wl["taps"], wl["IF"], "FOLD_ANTI", coeffs=h, n_channels=C, layout=wl["layout"], device=local)
        launches_per_step = 2
    elif wl["kind"] == "polydec":
        h = rng.integers(-32768, 32767, size=wl["taps"] * wl["df"], endpoint=True).astype(np.int16)
        f = E.ac_poly_dec(Q15, Q15, ACC40, ACC40, wl["taps"], wl["df"], coeffs=h, n_channels=C, layout=wl["layout"], device=local)
        launches_per_step = 2
    elif wl["kind"] == "cicfir":
        h = rng.integers(-32768, 32767, size=wl["taps"], endpoint=True).astype(np.int16)
        f = E.cic_intr_fir_cascade(Q15, wl["mid"], wl["R"], wl["M"], wl["N"], ACC40, Q15, ACC40, wl["taps"], "SHIFT_REG",
                                   coeffs=h, n_channels=C, layout=wl["layout"], device=local)
        launches_per_step = 2
    else:
        cls = E.ac_cic_dec_full if wl["mode"] == "dec" else E.ac_cic_intr_full
        f = cls(Q15, wl["out"], wl["R"], wl["M"], wl["N"], n_channels=C, layout=wl["layout"], device=local)
        launches_per_step = 2
    units_per_step = n if wl["unit_is_iq"] else n * C   # IQ pairs, or real samples over all local channels

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    y = f.run(x)
    path = f.path
    up = wl.get("R", 1) if (wl.get("mode") == "intr" or wl["kind"] == "cicfir") else wl.get("IF", 1)
    ybuf = torch.empty(max(y.numel(), C * n * up), dtype=y.dtype, device="cuda")
    del y
    for _ in range(args.warmup):
        y = f.run(x, out=ybuf)
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        y = f.run(x, out=ybuf)   # inputs + outputs per step (>= 5 GiB) exceed the 126 MB L2 many times over
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms / args.steps
    value = units_per_step * world / (ms_per_step * 1e-3) / 1e6
    out_bytes = y.numel() * y.element_size()
    in_bytes = x.numel() * x.element_size()
    del y

    # ---- end to end through the C-ABI host-buffer call (pinned host memory, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        n2 = min(n, 1 << 27 if wl["kind"] == "fir" else 1 << 28)
        shape2 = (n2, C) if il else ((C, n2) if C > 1 else (n2,))
        xh = torch.empty(shape2, dtype=x.dtype).pin_memory()
        xh.copy_(x[:n2] if (il or C == 1) else x[:, :n2])
        xn = xh.numpy()
        lib = E.load()
        import ctypes as ct
        if wl["kind"] == "fir":
            yh = torch.empty(n2 * C, dtype=torch.int64).pin_memory()
            call = lambda: lib.b2d_fir_run(f._h, xn.ctypes.data, n2, yh.data_ptr(), None)
        elif wl["kind"] == "intgdump":
            tok2 = np.full(n2 // (wl["chn"] * wl["nsamp"]), wl["nsamp"], dtype=np.uint32)
            yh = torch.empty(tok2.size * wl["chn"], dtype=torch.int32).pin_memory()
            call = lambda: lib.b2d_intgdump_run(f._h, xn.ctypes.data, n2, tok2.ctypes.data, tok2.size, yh.data_ptr(), ct.byref(no))
        elif wl["kind"] == "polyintr":
            yh = torch.empty(lib.b2d_polyintr_max_out(f._h, n2) * C, dtype=torch.int64).pin_memory()
            call = lambda: lib.b2d_polyintr_run(f._h, xn.ctypes.data, n2, yh.data_ptr(), ct.byref(no))
        elif wl["kind"] == "polydec":
            yh = torch.empty(lib.b2d_polydec_max_out(f._h, n2) * C, dtype=torch.int64).pin_memory()
            call = lambda: lib.b2d_polydec_run(f._h, xn.ctypes.data, n2, yh.data_ptr(), ct.byref(no))
        elif wl["kind"] == "cicfir":
            yh = torch.empty(lib.b2d_cicfir_max_out(f._h, n2) * C, dtype=torch.int64).pin_memory()
            call = lambda: lib.b2d_cicfir_run(f._h, xn.ctypes.data, n2, yh.data_ptr(), ct.byref(no))
        else:
            cap = lib.b2d_cic_max_out(f._h, n2)
            yh = torch.empty(cap * C, dtype=torch.int32).pin_memory()
            call = lambda: lib.b2d_cic_run(f._h, xn.ctypes.data, n2, yh.data_ptr(), ct.byref(no))
        no = ct.c_size_t(n2)
        for _ in range(2):
            assert call() == 0, lib.b2d_last_error()
        sync_all()
        t0 = time.perf_counter()
        k2 = max(3, min(args.steps, 5))
        for _ in range(k2):
            assert call() == 0, lib.b2d_last_error()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        u2 = n2 if wl["unit_is_iq"] else n2 * C
        e2e = {"value": u2 * world * k2 / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(xh.numel() * xh.element_size()),
               "d2h_bytes_per_step": int(no.value * C * yh.element_size()), "steps": k2,
               "api": "b2d_*_run (C-ABI, pinned host buffers, 3-slot copy/compute pipeline)",
               "host_affinity": "NVML ideal CPUs of the GPU" if orig_affinity else "unchanged",
               "host_numa_alloc": os.environ.get("B2D_HOST_NUMA") == "1",
               "samples_per_step": u2}

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = wl["bytes_per_unit"] * units_per_step
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp):
            t = json.load(open(tp)).get(args.workload)
            if t:   # measured DRAM bytes per unit (one ncu --set full capture) scaled to this launch's units
                traffic = t["dram_bytes_per_unit"] * units_per_step
                traffic_src = f"{t['source']}: {t['dram_bytes']} B measured at {t['capture_units']} units/launch, scaled"
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": path,
                "algorithmic_bytes_per_launch": alg_bytes, "actual_io_bytes_per_launch": in_bytes + out_bytes}
        if wl["macs_per_unit"] and path == "fir_q15":
            tmacs = wl["macs_per_unit"] * units_per_step / (ms_per_step * 1e-3) / 1e12
            # IDP.2A issue ceiling measured by tools/ubench_pipes.cu: 64 lanes/clk/SM, 2 16b x 8b products per lane-op,
            # 2 byte planes per 16 x 16 MAC -> 64 MAC/clk/SM
            sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
            roof["int_pipe"] = {"achieved_tmac_s": tmacs, "ceiling_tmac_s": 148 * 64 * sm_mhz * 1e6 / 1e12,
                                "frac": tmacs / (148 * 64 * sm_mhz * 1e6 / 1e12),
                                "note": "CUDA-core IDP.2A issue ceiling at the sampled SM clock (tensor cores excluded by the north star)"}
        line = {"metric": "Msamples/s (16b IQ, 256-tap FIR)" if args.workload == "fir256" else f"Msamples/s ({args.workload})",
                "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fir": "s16 x s16 -> s64 (exact integer, ac_fixed<40,8> wrap)", "cicfir": "s16 x s24 -> s64 (exact integer, ac_fixed<40,8> wrap)",
                          "polydec": "s16 x s16 -> s64 (exact integer, ac_fixed<40,8> wrap)",
                          "polyintr": "s16 x s16 -> s64 (exact integer, ac_fixed<40,8> wrap)",
                          "intgdump": "s16 -> s64 (exact integer sum, ac_fixed<32,17> wrap)",
                          "cic": "s16 -> u32 (modular integrate / comb)"}[wl["kind"]],
                "data": "synthetic",
                "config": {"workload": wl["name"], "samples_per_step_per_gpu": units_per_step, "kernel_path": path,
                           "l2": "inputs per step exceed L2 (>= 0.5 GiB vs 126 MB); no flush needed",
                           "parallelism": f"channels sharded over {world} GPU(s), one ncclBroadcast of the coefficient set at load()"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "roofline": roof}
        if world == 1 and not args.no_cpu:
            if orig_affinity:
                os.sched_setaffinity(0, orig_affinity)
            cb = cpu_reference(wl)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(line)
    f.close()
    if comm:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
