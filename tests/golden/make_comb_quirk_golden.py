"""tests/golden/make_comb_quirk_golden.py -- TEST INFRASTRUCTURE; run in the dev container (needs /root/reference).

Outputs of the UNMODIFIED reference CIC classes for differential delays M > 2, where the comb's ascending shift loop
(ac_cic_full_core.h:247-251) makes the effective delay 2.  oracle/_ref has no such instantiation (its table stops at
M = 2, like the reference's own tests), so the templates are compiled here for exactly these configurations, the same
way tests/test_oracle_fuzz.py does, and the results are committed as tests/golden/cic_comb_quirk.npz: that file is
what pins Oracle B and the CUDA engine to the real reference on the GPU box.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402
import test_oracle_fuzz as F  # noqa: E402

Q15 = (16, 1, True, "AC_TRN", "AC_WRAP")


def lossless(mode, fi, R, M, N):
    W = O.cic_int_width(mode, fi, R, M, N)
    return (W, W - (fi[0] - fi[1]), True, "AC_TRN", "AC_WRAP")


def configs():
    out = []
    for mode in ("dec", "intr"):
        for R, M, N in ((4, 3, 2), (8, 3, 4), (2, 5, 3), (7, 4, 1), (3, 3, 5)):
            out.append((mode, R, M, N, Q15, lossless(mode, Q15, R, M, N)))
        out.append((mode, 4, 3, 3, Q15, (16, 1, True, "AC_RND", "AC_SAT")))              # narrowed, saturating output
        out.append((mode, 5, 3, 2, (10, 2, False, "AC_TRN", "AC_WRAP"), (24, 12, True, "AC_TRN", "AC_WRAP")))
    return out


def main():
    cfgs = configs()
    dec = [c for c in cfgs if c[0] == "dec"]
    intr = [c for c in cfgs if c[0] == "intr"]
    incs = {"cfgs_cic_dec.inc": "".join(f"X({i}, {R}, {M}, {N}, {F.cfmt(fi)}, {F.cfmt(fo)})\n" for i, (_, R, M, N, fi, fo) in enumerate(dec)),
            "cfgs_cic_intr.inc": "".join(f"X({i}, {R}, {M}, {N}, {F.cfmt(fi)}, {F.cfmt(fo)})\n" for i, (_, R, M, N, fi, fo) in enumerate(intr))}
    tmp = tempfile.mkdtemp(prefix="combquirk")
    L = F.compile_driver(tmp, incs, [("ref_driver_cic.cpp", ["-DACREF_CIC_DEC"]), ("ref_driver_cic.cpp", ["-DACREF_CIC_INTR"])], "libq.so")
    import ctypes as C
    for fn in (L.acref_cic_dec_create, L.acref_cic_intr_create):
        fn.restype = C.c_void_p
        fn.argtypes = [C.c_int]
    L.acref_cic_run.restype = C.c_long
    L.acref_cic_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64)]
    L.acref_cic_destroy.argtypes = [C.c_void_p]
    rng = np.random.default_rng(20260101)
    store = {"n": np.array([len(cfgs)])}
    for k, (mode, R, M, N, fi, fo) in enumerate(cfgs):
        i = (dec if mode == "dec" else intr).index((mode, R, M, N, fi, fo))
        n = 900 if mode == "dec" else 150
        x = np.ascontiguousarray(O.rand_raw(rng, fi, n), dtype=np.int64)
        x[:3] = [O.rand_raw(rng, fi, 1, "min")[0], O.rand_raw(rng, fi, 1, "max")[0], O.rand_raw(rng, fi, 1, "min")[0]]
        h = (L.acref_cic_dec_create if mode == "dec" else L.acref_cic_intr_create)(i)
        parts = []
        for lo, hi in ((0, 1), (1, 10), (10, 13), (13, n)):
            seg = np.ascontiguousarray(x[lo:hi])
            buf = np.empty(seg.size * R + R + 8, dtype=np.int64)
            m = L.acref_cic_run(h, F.p64(seg), seg.size, F.p64(buf))
            parts.append(buf[:m].copy())
        L.acref_cic_destroy(h)
        store[f"c{k}_mode"] = np.array([0 if mode == "dec" else 1])
        store[f"c{k}_rmn"] = np.array([R, M, N])
        store[f"c{k}_fin"] = np.array([fi[0], fi[1], int(fi[2]), O.Q_MODES.index(fi[3]), O.O_MODES.index(fi[4])])
        store[f"c{k}_fout"] = np.array([fo[0], fo[1], int(fo[2]), O.Q_MODES.index(fo[3]), O.O_MODES.index(fo[4])])
        store[f"c{k}_x"] = x
        store[f"c{k}_y"] = np.concatenate(parts)
        store[f"c{k}_counts"] = np.array([p.size for p in parts], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "cic_comb_quirk.npz"), **store)
    print("cic_comb_quirk.npz:", len(cfgs), "configurations")


if __name__ == "__main__":
    main()
