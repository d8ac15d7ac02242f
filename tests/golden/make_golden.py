"""tests/golden/make_golden.py -- regenerate the committed fixtures in tests/golden/.

Run in the dev container (needs /root/reference and oracle/_ref/libacdsp_ref.so):
    python tests/golden/make_golden.py

Fixtures (all raw two's-complement integers, int64):
  cic_dec_golden.npz / cic_intr_golden.npz
      the reference's own MATLAB-generated bit-exact vectors
      (tests/ac_cic_{dec,intr}_full_{input,ref}.txt, values are exact multiples of 2^-16),
      converted to raw integers, with the bench's framing applied
      (rtest_ac_cic_dec_full.cpp:78-85 prepends one 0; rtest_ac_cic_intr_full.cpp:88-99 uses
      the first 1000 inputs and skips the first N=5 reference values).
  fir_bench_{const,load,prog}.npz
      the reference FIR benches' stimulus (two-tone, rtest_ac_fir_const_coeffs.cpp:126-151),
      coefficient files, the MATLAB double reference, and the output of the UNMODIFIED
      reference class (Oracle A) on that stimulus, plus the SQNR it obtains.
  rs_outputs.npz
      outputs of the UNMODIFIED reference class ac_fir_reg_share (oracle/ref_driver_rs.cpp) for every configuration
      in oracle/ref_configs.RS_CONFIGS: samples, coefficient RAM image, outputs of three run batches, final
      ac_firProgCoeffs_delay_line value; and of the UNMODIFIED ac_poly_dec (oracle/ref_driver_pd.cpp) for every
      configuration in PD_CONFIGS (samples, phase-ordered coefficients, outputs of three run batches), of the
      UNMODIFIED ac_poly_intr (oracle/ref_driver_pi.cpp) for every configuration in PI_CONFIGS (two coefficient /
      control sets, the second loaded half way through the stream), of ac_intg_dump (ID_CONFIGS) and of ac_mv_avg
      (MV_CONFIGS; the unmodified class over the RESTATED ac_window_1d_flag of oracle/ac_shim/ac_window.h: parity unpinned).
  ref_outputs.npz
      outputs of the UNMODIFIED reference classes (Oracle A) on seeded random inputs for every
      configuration in oracle/ref_configs.py x ftype (FIR) and every CIC configuration, fed in
      several run() calls.  The GPU box has no /root/reference; these pin the CUDA path to the
      real reference there.
"""
import math
import os
import sys
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle import ref_configs as rc  # noqa: E402

REF = os.environ.get("AC_DSP_REF", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
SEED = 20260101


def read_exact(path, F):
    """Decimal text -> raw integers value*2^F, checked to be exact."""
    vals = []
    for tok in open(path).read().replace(",", " ").split():
        fr = Fraction(tok) * (1 << F)
        assert fr.denominator == 1, (path, tok)
        vals.append(int(fr))
    return np.array(vals, dtype=np.int64)


def read_doubles(path):
    return np.array([float(t) for t in open(path).read().replace(",", " ").split()], dtype=np.float64)


def bench_stimulus(fin, n=1024):
    """rtest_ac_fir_const_coeffs.cpp:126-151: two tones, normalised to the type maximum, AC_TRN."""
    W, I, S, _, _ = O.normfmt(fin)
    F = W - I
    tmax = ((1 << (W - 1)) - 1) / float(1 << F)
    pi = 3.14159265358979323846
    tones = [math.sin(2 * pi * 25 * i / 500.0) + math.sin(2 * pi * 150 * i / 500.0) for i in range(n)]
    amax = max(abs(t) for t in tones)
    raw = [math.floor(((t / amax) * tmax) * float(1 << F)) for t in tones]
    return np.array(raw, dtype=np.int64)


def sqnr(out_raw, F, ref):
    d = out_raw.astype(np.float64) / float(1 << F)
    return 10 * math.log10(np.sum(ref * ref) / np.sum((d - ref) ** 2))


def main():
    t = REF + "/tests/"
    # ---- CIC goldens
    din = read_exact(t + "ac_cic_dec_full_input.txt", 16)
    dref = read_exact(t + "ac_cic_dec_full_ref.txt", 16)
    din = np.concatenate([[0], din]).astype(np.int64)
    a = O.CicA("dec", (32, 16), (48, 32), 7, 2, 4)
    got = a.run(din)
    assert np.array_equal(got[: dref.size], dref), "reference class does not reproduce its own golden vector"
    np.savez_compressed(OUT + "/cic_dec_golden.npz", x=din, ref=dref, R=7, M=2, N=4, fin=[32, 16, 1], fout=[48, 32, 1])
    iin = read_exact(t + "ac_cic_intr_full_input.txt", 16)[:1000]
    iref = read_exact(t + "ac_cic_intr_full_ref.txt", 16)[5:]
    a = O.CicA("intr", (32, 16), (49, 33), 7, 2, 5)
    got = a.run(iin)
    assert got.size == 6990 and np.array_equal(got, iref[: got.size])
    np.savez_compressed(OUT + "/cic_intr_golden.npz", x=iin, ref=iref[: got.size], R=7, M=2, N=5, fin=[32, 16, 1], fout=[49, 33, 1])
    print("cic goldens:", din.size, "->", dref.size, ";", iin.size, "->", got.size)

    # ---- FIR benches
    benches = {
        "const": ("bench_const", 29, "ac_fir_const_coeffs", 84.2385),
        "load": ("bench_load", 27, "ac_fir_load_coeffs", 89.5576),
        "prog": ("bench_prog", 27, "ac_fir_prog_coeffs", 89.5576),
    }
    fmts = {name: (fi, fc, fa, fo) for name, fi, fc, fa, fo, _ in rc.FIR_FORMATS}
    for cls, (fname, taps, stem, want) in benches.items():
        fi, fc, fa, fo = fmts[fname]
        Fc = fc[0] - fc[1]
        cd = read_doubles(t + stem + "_cfg.txt")
        assert cd.size == taps
        craw = np.array([math.floor(c * (1 << Fc)) for c in cd], dtype=np.int64)
        x = bench_stimulus(fi)
        ref = read_doubles(t + stem + "_ref.txt")
        f = O.FirA(cls, fi, fc, fa, fo, taps, "FOLD_ODD")
        f.load(craw)
        y = f.run(x)
        s = sqnr(y, fo[0] - fo[1], ref[: y.size])
        print(f"fir bench {cls}: SQNR {s:.4f} dB (reference bench prints {want})")
        assert abs(s - want) < 5e-4
        np.savez_compressed(OUT + f"/fir_bench_{cls}.npz", x=x, coeffs=craw, y=y, ref_double=ref, sqnr=s,
                            fin=fi[:3], fcoeff=fc[:3], facc=fa[:3], fout=fo[:3], taps=taps)

    # ---- reference outputs on seeded random inputs
    rng = np.random.default_rng(SEED)
    store = {}
    for cid, name, fi, fc, fa, fo, taps in rc.fir_configs():
        n = 3 * taps + 40 if taps >= 256 else 320
        x = O.rand_raw(rng, fi, n)
        c = O.rand_raw(rng, fc, taps)
        csym = c.copy()
        csym[taps - (taps // 2):] = c[: taps // 2][::-1]          # symmetric set for the FOLD types
        store[f"fir{cid}_x"] = x
        store[f"fir{cid}_c"] = c
        store[f"fir{cid}_csym"] = csym
        for ft in O.FTYPES[:6]:
            cls = O.FIR_CLASSES[(cid + O.FTYPES.index(ft)) % 3]
            f = O.FirA(cls, fi, fc, fa, fo, taps, ft)
            f.load(csym if ft.startswith("FOLD") else c)
            y = np.concatenate([f.run(x[:5]), f.run(x[5:6]), f.run(x[6:])])
            store[f"fir{cid}_{ft}_y"] = y
    for cid, (mode, R, M, N, fi, fo) in enumerate(rc.CIC_CONFIGS):
        n = 260 if mode == "dec" else 90
        x = O.rand_raw(rng, fi, n)
        f = O.CicA(mode, fi, fo, R, M, N)
        parts = [f.run(x[:1]), f.run(x[1:10]), f.run(x[10:13]), f.run(x[13:])]
        store[f"cic{cid}_x"] = x
        store[f"cic{cid}_y"] = np.concatenate(parts)
        store[f"cic{cid}_counts"] = np.array([p.size for p in parts], dtype=np.int64)
    # ---- ac_fir_reg_share: the real reference class on seeded inputs, separate file
    rs = {}
    rng2 = np.random.default_rng(SEED + 1)
    for cid, cfg in enumerate(rc.RS_CONFIGS):
        N, fi, fo, fc, fa, mww, bs, bo, ft = cfg
        f = O.RsA(cid)
        x = O.rand_raw(rng2, fi, 4 * N + 50)
        ram = O.rand_raw(rng2, fc, f.ram_words)
        rs[f"rs{cid}_x"], rs[f"rs{cid}_ram"] = x, ram
        rs[f"rs{cid}_y"] = np.concatenate([f.run(x[:7], ram), f.run(x[7:8], ram), f.run(x[8:], ram)])
        rs[f"rs{cid}_dl"] = np.array([f.delay_out()], dtype=np.int64)
    for cid, (fi, fc, fa, fo, nt, df) in enumerate(rc.PD_CONFIGS):      # ac_poly_dec (row N2), same file
        f = O.PdA(cid)
        x = O.rand_raw(rng2, fi, 6 * nt * df + 37)
        c = O.rand_raw(rng2, fc, nt * df)
        f.load(c)
        rs[f"pd{cid}_x"], rs[f"pd{cid}_c"] = x, c
        rs[f"pd{cid}_y"] = np.concatenate([f.run(x[:1]), f.run(x[1:df + 2]), f.run(x[df + 2:])])
    for cid, (fi, fa, fo, NS, CHN) in enumerate(rc.ID_CONFIGS):         # ac_intg_dump (row N4), same file
        f = O.IdA(cid)
        ys, xs, toks = [], [], []
        for call in range(3):                                            # three run() calls: the sums carry across them
            ns = rng2.integers(1, NS, size=6, endpoint=True)
            ns[1], ns[4] = NS + 2, 0                                     # frames that do not dump
            if call == 2:
                ns[5] = NS + 1                                           # ... also at the very end of a call
            x = O.rand_raw(rng2, fi, sum(O.id_frame_samples(v, NS, CHN) for v in ns))
            ys.append(f.run(x, ns)); xs.append(x); toks.append(ns)
        rs[f"id{cid}_x"], rs[f"id{cid}_ns"], rs[f"id{cid}_y"] = np.concatenate(xs), np.concatenate(toks), np.concatenate(ys)
        rs[f"id{cid}_xlen"] = np.array([v.size for v in xs], dtype=np.int64)
    rng3 = np.random.default_rng(SEED + 4)                              # own stream: adding rows here leaves the others alone
    for cid, (maxs, taps, wt, fi, fo, fa, fc) in enumerate(rc.MV_CONFIGS):   # ac_mv_avg (row N4): unmodified class over the RESTATED window
        c = O.rand_raw(rng3, fc, taps)
        f = O.MvA(cid, c)
        ns1, ns2 = min(maxs, 3 * taps + 5), taps                        # two run() calls: three bursts, then two of the shortest legal length
        x1, x2 = O.rand_raw(rng3, fi, 3 * ns1), O.rand_raw(rng3, fi, 2 * ns2)
        x1[:2] = [O.rand_raw(rng3, fi, 1, "min")[0], O.rand_raw(rng3, fi, 1, "max")[0]]
        rs[f"mv{cid}_c"], rs[f"mv{cid}_x1"], rs[f"mv{cid}_x2"] = c, x1, x2
        rs[f"mv{cid}_ns"] = np.array([ns1, ns2], dtype=np.int64)
        rs[f"mv{cid}_y1"], rs[f"mv{cid}_y2"] = f.run(x1, ns1), f.run(x2, ns2)
    for cid, cfg in enumerate(rc.PI_CONFIGS):                           # ac_poly_intr (row N2), same file
        fi, fc, fa, fo, nt, IF, ft = cfg
        f = O.PiA(cid)
        x = O.rand_raw(rng2, fi, 5 * nt + 41)
        x[nt + 3] = -(1 << (fi[0] - 1)) if fi[2] else (1 << fi[0]) - 1      # the sample whose negation does not fit IN_TYPE
        c1, c2 = O.rand_raw(rng2, fc, f.coeffsz), O.rand_raw(rng2, fc, f.coeffsz)
        sign = rng2.integers(0, 2, size=IF)
        corr = np.arange(IF)
        if IF >= 2:                                                      # symmetric pairs (j, IF-1-j), as the technique pairs them
            corr = IF - 1 - corr
            corr[IF // 2:] = np.arange(IF)[IF // 2:] if cid % 2 else corr[IF // 2:]
        half = x.size // 2
        f.load(c1, sign, corr)
        y1 = np.concatenate([f.run(x[:1]), f.run(x[1:9]), f.run(x[9:half])])
        f.load(c2, 1 - sign, corr[::-1].copy())                          # reload between steps: the parked accumulators keep their values
        y2 = f.run(x[half:])
        rs[f"pi{cid}_x"], rs[f"pi{cid}_c1"], rs[f"pi{cid}_c2"] = x, c1, c2
        rs[f"pi{cid}_sign"], rs[f"pi{cid}_corr"] = sign.astype(np.int64), corr.astype(np.int64)
        rs[f"pi{cid}_y"] = np.concatenate([y1, y2])
        rs[f"pi{cid}_half"] = np.array([half, y1.size], dtype=np.int64)
    np.savez_compressed(OUT + "/rs_outputs.npz", **rs)
    print("rs_outputs:", len(rs), "arrays")
    np.savez_compressed(OUT + "/ref_outputs.npz", **store)
    print("ref_outputs:", len(store), "arrays,", os.path.getsize(OUT + "/ref_outputs.npz") // 1024, "KiB")


if __name__ == "__main__":
    main()
