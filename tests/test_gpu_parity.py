"""GPU suite (-m gpu): the CUDA engine, called through its C-ABI, against
  * the outputs of the UNMODIFIED reference classes (tests/golden/ref_outputs.npz, the two CIC golden vectors,
    the three FIR bench vectors) -- bit-exact, and
  * the CPU oracle (oracle/oracle_b.c) on seeded random inputs -- bit-exact,
  * size-independent properties at BASELINE.json's full sizes (random windows re-derived by the oracle).
Nothing here reads /root/reference.
"""
import math
import os
import zlib

import numpy as np
import pytest

from conftest import golden
from oracle import ref_configs as rc

pytestmark = pytest.mark.gpu

Q15, ACC40 = (16, 1), (40, 8)
KINDS = ["const", "load", "prog"]


def make_fir(E, kind, fi, fc, fa, fo, taps, ft, coeffs, **kw):
    if kind == "const":
        return E.ac_fir_const_coeffs(fi, fo, fc, fa, taps, ft, coeffs, **kw)
    if kind == "load":
        f = E.ac_fir_load_coeffs(fi, fo, fc, fa, taps, ft, **kw)
        f.run(None, coeffs, True)
        return f
    f = E.ac_fir_prog_coeffs(fi, fo, fc, fa, taps, ft, **kw)
    f.load(coeffs)
    return f


@pytest.fixture(params=["auto", "generic", "wide"])
def path(request, monkeypatch):
    """Every golden case runs through the kernel family the engine would pick, through the generic kernels, and with
    the 16-bit DP2A path disabled (so the IMAD.WIDE path also sees the 16-bit formats)."""
    if request.param == "generic":
        monkeypatch.setenv("B2D_FORCE_GENERIC", "1")
    elif request.param == "wide":
        monkeypatch.setenv("B2D_FORCE_GENERIC", "2")
    else:
        monkeypatch.delenv("B2D_FORCE_GENERIC", raising=False)
    return request.param


@pytest.fixture(autouse=True)
def _dp2a_unless_asked(monkeypatch):
    """This module pins the DP2A kernel (fir_q15) unless a test asks for the `ovs` fixture: the overlap-save evaluation
    would otherwise take every long call of a 96..2049-tap filter (tests/test_fir_ovs.py covers its own selection rules)."""
    monkeypatch.setenv("B2D_FIR_OVS", "0")


@pytest.fixture(params=["q15", "ovs"])
def ovs(request, monkeypatch):
    """Long 16-bit filters through both evaluations: the DP2A kernel, and overlap-save forced for every call length
    (B2D_FIR_OVS=2) with the residual monitor on."""
    monkeypatch.setenv("B2D_FIR_OVS", "2" if request.param == "ovs" else "0")
    monkeypatch.setenv("B2D_OVS_RESID", "1")
    return request.param


@pytest.fixture(params=["q15", "auto"])
def ovs_auto(request, monkeypatch):
    """Full-size calls: the DP2A kernel, and the engine's own choice (overlap-save for long calls of long filters)."""
    if request.param == "auto":
        monkeypatch.delenv("B2D_FIR_OVS", raising=False)
    return request.param


def check_ovs(f, ovs, taps):
    """Path name and exactness margin of a q15-format filter under the `ovs` fixture."""
    bound, resid = f.ovs_margin()
    if ovs == "ovs" and 96 <= taps <= 2049 and bound < 0.49:
        assert f.path == "fir_ovs"
        assert 0 <= resid < 0.01, (bound, resid)      # far inside the a-priori bound
    else:
        assert f.path == "fir_q15"


@pytest.mark.parametrize("cfg", rc.fir_configs(), ids=lambda c: f"{c[1]}-{c[6]}")
def test_fir_vs_reference_outputs(engine, ref_outputs, cfg, path):
    cid, _name, fi, fc, fa, fo, taps = cfg
    x = ref_outputs[f"fir{cid}_x"]
    for k, ft in enumerate(engine.FTYPES[:6]):
        c = ref_outputs[f"fir{cid}_csym" if ft.startswith("FOLD") else f"fir{cid}_c"]
        f = make_fir(engine, KINDS[(cid + k) % 3], fi, fc, fa, fo, taps, ft, c)
        if path == "generic":
            assert f.path == "fir_generic"
        y = np.concatenate([f.run(x[:5]), f.run(x[5:6]), f.run(x[6:])]).astype(np.int64)
        assert np.array_equal(y, ref_outputs[f"fir{cid}_{ft}_y"]), (cfg, ft, f.path)


def test_q15_path_is_taken_for_the_baseline_formats(engine):
    h = np.ones(256, dtype=np.int16)
    for ft in ("SHIFT_REG", "ROTATE_SHIFT", "C_BUFF", "TRANSPOSED", "FOLD_EVEN", "FOLD_ODD"):
        assert make_fir(engine, "load", Q15, Q15, ACC40, ACC40, 256, ft, h).path == "fir_q15"
    assert make_fir(engine, "load", (20, 5), Q15, ACC40, ACC40, 63, "SHIFT_REG", h[:63]).path == "fir_q24"
    assert make_fir(engine, "load", (32, 16), Q15, (64, 32), (64, 32), 27, "SHIFT_REG", h[:27]).path == "fir_wide"
    # order-dependent accumulators (saturation, sign-dependent rounding) only exist on the generic kernel
    assert make_fir(engine, "load", Q15, Q15, (24, 4, True, "AC_TRN", "AC_SAT"), Q15, 16, "SHIFT_REG", h[:16]).path == "fir_generic"
    assert make_fir(engine, "load", Q15, Q15, (24, 4, True, "AC_TRN_ZERO"), Q15, 16, "SHIFT_REG", h[:16]).path == "fir_generic"
    assert make_fir(engine, "load", Q15, Q15, (24, 4), Q15, 16, "SHIFT_REG", h[:16]).path == "fir_wide"  # per-tap truncation


def test_cic_vs_reference_outputs(engine, ref_outputs, path):
    for cid, (mode, R, M, N, fi, fo) in enumerate(rc.CIC_CONFIGS):
        x = ref_outputs[f"cic{cid}_x"]
        cls = engine.ac_cic_dec_full if mode == "dec" else engine.ac_cic_intr_full
        f = cls(fi, fo, R, M, N)
        parts = [f.run(x[:1]), f.run(x[1:10]), f.run(x[10:13]), f.run(x[13:])]
        assert [p.size for p in parts] == list(ref_outputs[f"cic{cid}_counts"]), (mode, R, M, N)
        assert np.array_equal(np.concatenate(parts).astype(np.int64), ref_outputs[f"cic{cid}_y"]), (mode, R, M, N, f.path)


def test_cic_golden_vectors(engine, path):
    g = golden("cic_dec_golden.npz")
    y = engine.ac_cic_dec_full((32, 16), (48, 32), 7, 2, 4).run(g["x"])
    assert y.size == 1430 and np.array_equal(y[:1429], g["ref"])
    g = golden("cic_intr_golden.npz")
    y = engine.ac_cic_intr_full((32, 16), (49, 33), 7, 2, 5).run(g["x"])
    assert y.size == 6990 and np.array_equal(y, g["ref"])


@pytest.mark.parametrize("cls", KINDS)
def test_fir_reference_benches(engine, cls, path):
    g = golden(f"fir_bench_{cls}.npz")
    fi, fc, fa, fo = (tuple(int(v) for v in g[k]) for k in ("fin", "fcoeff", "facc", "fout"))
    f = make_fir(engine, cls, fi, fc, fa, fo, int(g["taps"]), "FOLD_ODD", g["coeffs"])
    if cls == "prog":   # the bench calls run() once per sample (rtest_ac_fir_prog_coeffs.cpp:109-113)
        y = np.concatenate([f.run(g["x"][i:i + 1], g["coeffs"]) for i in range(64)] + [f.run(g["x"][64:], g["coeffs"])])
    else:
        y = f.run(g["x"])
    assert np.array_equal(y.astype(np.int64), g["y"])
    ref = g["ref_double"][: y.size]
    sqnr = 10 * math.log10(np.sum(ref * ref) / np.sum((y / float(1 << (fo[0] - fo[1])) - ref) ** 2))
    assert abs(sqnr - float(g["sqnr"])) < 1e-9 and sqnr >= 60.0


# ------------------------------------------------------------------------ oracle-differential, FIR
def oracle_fir(O, fi, fc, fa, fo, taps, ft, coeffs, x):
    f = O.FirB(fi, fc, fa, fo, taps, ft)
    f.load(coeffs)
    return f.run(x)


@pytest.mark.parametrize("taps,n", [(1, 100), (2, 1000), (16, 4096), (16, 1 << 20), (27, 5000), (63, 9999), (255, 8192),
                                    (256, 4095), (256, 70001), (1024, 20000), (2048, 9000)])
def test_fir_q15_random(engine, oracle, taps, n, ovs):
    rng = np.random.default_rng(taps * 7919 + n)
    x = oracle.rand_raw(rng, Q15, n)
    h = oracle.rand_raw(rng, Q15, taps)
    f = make_fir(engine, "load", Q15, Q15, ACC40, ACC40, taps, "SHIFT_REG", h)
    y = f.run(x.astype(np.int16))
    assert np.array_equal(y.astype(np.int64), oracle_fir(oracle, Q15, Q15, ACC40, ACC40, taps, "SHIFT_REG", h, x))
    check_ovs(f, ovs, taps)


@pytest.mark.parametrize("kind", ["min", "max", "alt"])
def test_fir_q15_extremes_and_wrap(engine, oracle, kind, ovs):
    """all-(-1.0) x all-(-1.0) over 256 taps reaches 2^40 and wraps to 0 in <40,8> (DERIVED KAT)."""
    rng = np.random.default_rng(3)
    for taps in (256, 1024):
        x = oracle.rand_raw(rng, Q15, 3000, kind)
        h = oracle.rand_raw(rng, Q15, taps, "min" if kind != "max" else "max")
        y = make_fir(engine, "const", Q15, Q15, ACC40, ACC40, taps, "C_BUFF", h).run(x)
        want = oracle_fir(oracle, Q15, Q15, ACC40, ACC40, taps, "C_BUFF", h, x)
        assert np.array_equal(y.astype(np.int64), want)
        if kind == "min" and taps == 256:
            assert y[255] == 0 and y[254] == (255 << 32) - (1 << 40)


def test_fir_q15_iq_interleaved_and_planar(engine, oracle, ovs):
    rng = np.random.default_rng(11)
    n, taps = 40000 + 3, 256
    x = rng.integers(-32768, 32767, size=(n, 2), endpoint=True).astype(np.int16)
    h = oracle.rand_raw(rng, Q15, taps)
    want = [oracle_fir(oracle, Q15, Q15, ACC40, ACC40, taps, "SHIFT_REG", h, x[:, c]) for c in range(2)]
    f = make_fir(engine, "load", Q15, Q15, ACC40, ACC40, taps, "SHIFT_REG", h, n_channels=2, layout="interleaved")
    y = f.run(x)
    assert y.shape == (n, 2) and y.dtype == np.int64
    for c in range(2):
        assert np.array_equal(y[:, c], want[c])
    check_ovs(f, ovs, taps)
    f = make_fir(engine, "load", Q15, Q15, ACC40, ACC40, taps, "SHIFT_REG", h, n_channels=2, layout="planar")
    y = f.run(np.ascontiguousarray(x.T))
    for c in range(2):
        assert np.array_equal(y[c], want[c])
    check_ovs(f, ovs, taps)


def test_fir_per_channel_coefficients_and_layouts(engine, oracle, ovs):
    """cfg 4 shape (scaled down): independent channels, a coefficient set per channel, 1024 taps."""
    rng = np.random.default_rng(12)
    C, n, taps = 5, 6001, 1024
    x = rng.integers(-32768, 32767, size=(C, n), endpoint=True).astype(np.int16)
    hs = [oracle.rand_raw(rng, Q15, taps) for _ in range(C)]
    want = [oracle_fir(oracle, Q15, Q15, ACC40, ACC40, taps, "SHIFT_REG", hs[c], x[c]) for c in range(C)]
    for layout in ("planar", "interleaved"):
        f = engine.ac_fir_prog_coeffs(Q15, ACC40, Q15, ACC40, taps, "SHIFT_REG", n_channels=C, layout=layout)
        for c in range(C):
            f.load(hs[c], channel=c)
        xin = x if layout == "planar" else np.ascontiguousarray(x.T)
        y = f.run(xin)
        y = y if layout == "planar" else y.T
        for c in range(C):
            assert np.array_equal(y[c], want[c]), (layout, c)


@pytest.mark.parametrize("ft", ["SHIFT_REG", "FOLD_EVEN", "FOLD_ODD", "TRANSPOSED"])
def test_fir_chunk_split_invariance(engine, oracle, ft):
    """One run() vs arbitrary splits (1-sample calls, chunks < N_TAPS): the delay line is carried exactly."""
    rng = np.random.default_rng(13)
    taps, n = 63, 3000
    x = oracle.rand_raw(rng, Q15, n)
    h = oracle.rand_raw(rng, Q15, taps)
    want = oracle_fir(oracle, Q15, Q15, ACC40, ACC40, taps, ft, h, x)
    f = make_fir(engine, "load", Q15, Q15, ACC40, ACC40, taps, ft, h)
    cuts = [0, 1, 2, 10, 11, 70, 71, 500, 1999, n]
    y = np.concatenate([f.run(x[a:b].astype(np.int16)) for a, b in zip(cuts[:-1], cuts[1:])])
    assert np.array_equal(y.astype(np.int64), want)
    f.reset()
    assert np.array_equal(f.run(x.astype(np.int16)).astype(np.int64), want)


def test_fir_prog_coefficient_change_keeps_delay_line(engine, oracle):
    rng = np.random.default_rng(14)
    taps = 16
    x = oracle.rand_raw(rng, Q15, 400)
    h1, h2 = oracle.rand_raw(rng, Q15, taps), oracle.rand_raw(rng, Q15, taps)
    ob = oracle.FirB(Q15, Q15, ACC40, ACC40, taps, "SHIFT_REG")
    ob.load(h1)
    w1 = ob.run(x[:150])
    ob.load(h2)
    w2 = ob.run(x[150:])
    f = engine.ac_fir_prog_coeffs(Q15, ACC40, Q15, ACC40, taps, "SHIFT_REG")
    y = np.concatenate([f.run(x[:150].astype(np.int16), h1), f.run(x[150:].astype(np.int16), h2)])
    assert np.array_equal(y.astype(np.int64), np.concatenate([w1, w2]))


@pytest.mark.parametrize("fmts", [
    ((16, 1), (16, 1), (24, 4), (16, 1)),                                   # per-tap truncation, narrow output
    ((16, 1), (16, 1), (24, 4, True, "AC_RND"), (16, 1, True, "AC_RND")),    # rounding
    ((16, 1), (16, 1), (24, 4, True, "AC_TRN", "AC_SAT"), (16, 1, True, "AC_RND_CONV", "AC_SAT_SYM")),  # order-dependent
    ((16, 1), (16, 1), (30, 6, True, "AC_TRN_ZERO", "AC_SAT_ZERO"), (12, 1, True, "AC_RND_INF", "AC_SAT")),
    ((12, 0, False), (14, 2), (30, 6), (20, 4)),                             # unsigned input
    ((32, 16), (32, 16), (64, 32), (64, 32)),                                # bench_load formats
    ((28, 6), (23, 7), (64, 32), (64, 32)),                                  # bench_prog formats
    ((16, 1), (16, 1), (40, 8), (18, 2, True, "AC_RND", "AC_SAT")),          # q15 path with a converting epilogue
    ((16, 0, False), (16, 0, False), (48, 16, False), (48, 16, False)),      # unsigned q15 path
    ((16, 1), (12, 4, False), (36, 9), (36, 9)),                             # signed x unsigned coefficients
])
def test_fir_formats_random(engine, oracle, fmts):
    """Every ftype, including saturating / convergent-rounding accumulators (reference tap order matters there)."""
    fi, fc, fa, fo = fmts
    rng = np.random.default_rng(zlib.crc32(str(fmts).encode()))
    for taps in (7, 10, 33):
        x = oracle.rand_raw(rng, fi, 700)
        h = oracle.rand_raw(rng, fc, taps)
        h[taps - taps // 2:] = h[: taps // 2][::-1]
        for ft in engine.FTYPES[:6]:
            want = oracle_fir(oracle, fi, fc, fa, fo, taps, ft, h, x)
            f = make_fir(engine, "load", fi, fc, fa, fo, taps, ft, h)
            y = np.concatenate([f.run(x[:100]), f.run(x[100:])])
            assert np.array_equal(y.astype(np.int64), want), (fmts, taps, ft, f.path)


def test_fir_device_path_and_state(engine, oracle, ovs):
    import torch
    rng = np.random.default_rng(15)
    taps, n = 256, 50000
    x = oracle.rand_raw(rng, Q15, n)
    h = oracle.rand_raw(rng, Q15, taps)
    want = oracle_fir(oracle, Q15, Q15, ACC40, ACC40, taps, "SHIFT_REG", h, x)
    f = make_fir(engine, "load", Q15, Q15, ACC40, ACC40, taps, "SHIFT_REG", h)
    xd = torch.from_numpy(x.astype(np.int16)).cuda()
    y1 = f.run(xd[:20000])
    blob = f.get_state()
    y2 = f.run(xd[20000:])
    torch.cuda.synchronize()
    assert np.array_equal(torch.cat([y1, y2]).cpu().numpy(), want)
    g = make_fir(engine, "load", Q15, Q15, ACC40, ACC40, taps, "SHIFT_REG", h)
    g.set_state(blob)                       # checkpoint / resume
    assert np.array_equal(g.run(xd[20000:]).cpu().numpy(), want[20000:])
    check_ovs(g, ovs, taps)


def test_fir_call_sequence_errors(engine):
    f = engine.ac_fir_load_coeffs(Q15, ACC40, Q15, ACC40, 16)
    with pytest.raises(engine.B2dError) as e:
        f.run(np.zeros(10, dtype=np.int16))
    assert e.value.status == -6             # run() before coefficients
    f.run(None, np.ones(8), True)           # under-filled coefficient channel: token silently dropped (load:328)
    with pytest.raises(engine.B2dError):
        f.run(np.zeros(10, dtype=np.int16))
    c = engine.ac_fir_const_coeffs(Q15, ACC40, Q15, ACC40, 16, "SHIFT_REG", np.ones(16))
    with pytest.raises(engine.B2dError):
        c._load(np.ones(16))                # constant coefficients are bound once
    assert c.run(np.zeros(0, dtype=np.int16)).size == 0


# ------------------------------------------------------------------------ oracle-differential, CIC
CIC_CASES = [("dec", 8, 1, 4, Q15, (28, 13)), ("dec", 8, 2, 4, Q15, (32, 17)), ("dec", 7, 2, 4, (32, 16), (48, 32)),
             ("dec", 2, 1, 1, Q15, (17, 2)), ("dec", 16, 1, 5, Q15, (36, 21)), ("dec", 256, 1, 2, Q15, (32, 17)),
             ("dec", 8, 1, 4, Q15, (16, 1)), ("dec", 5, 1, 3, (10, 2, False), (24, 12)),
             ("intr", 4, 1, 3, Q15, (20, 5)), ("intr", 7, 2, 5, (32, 16), (49, 33)), ("intr", 2, 1, 5, Q15, (20, 5)),
             ("intr", 8, 2, 4, Q15, (29, 14)), ("intr", 16, 1, 1, Q15, (16, 1)), ("intr", 4, 2, 3, Q15, (12, 4, True, "AC_RND"))]


@pytest.mark.parametrize("case", CIC_CASES, ids=lambda c: f"{c[0]}-R{c[1]}M{c[2]}N{c[3]}-{c[4][0]}")
def test_cic_random_and_chunked(engine, oracle, case, path):
    mode, R, M, N, fi, fo = case
    rng = np.random.default_rng(R * 100 + M * 10 + N)
    n = 50000 if mode == "dec" else 9000
    x = oracle.rand_raw(rng, fi, n)
    want = oracle.CicB(mode, fi, fo, R, M, N).run(x)
    cls = engine.ac_cic_dec_full if mode == "dec" else engine.ac_cic_intr_full
    f = cls(fi, fo, R, M, N)
    y = f.run(x)
    assert np.array_equal(y.astype(np.int64), want), f.path
    f.reset()
    cuts = [0, 1, 2, 3, 9, 10, 10 + R, 11 + R, 500, 501, 7777, n]
    parts = [f.run(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate(parts).astype(np.int64), want)


def test_cic_multichannel_layouts_and_state(engine, oracle):
    rng = np.random.default_rng(21)
    C, n = 3, 30001
    x = rng.integers(-32768, 32767, size=(C, n), endpoint=True).astype(np.int16)
    for mode, R, M, N, fo in (("dec", 8, 1, 4, (28, 13)), ("intr", 4, 1, 3, (20, 5))):
        cls = engine.ac_cic_dec_full if mode == "dec" else engine.ac_cic_intr_full
        want = [oracle.CicB(mode, Q15, fo, R, M, N).run(x[c]) for c in range(C)]
        for layout in ("planar", "interleaved"):
            f = cls(Q15, fo, R, M, N, n_channels=C, layout=layout)
            xin = x if layout == "planar" else np.ascontiguousarray(x.T)
            half = (n // 2) | 1
            a = f.run(xin[:, :half] if layout == "planar" else xin[:half])
            blob = f.get_state()
            b = f.run(xin[:, half:] if layout == "planar" else xin[half:])
            for c in range(C):
                assert np.array_equal(np.concatenate([a[c], b[c]]).astype(np.int64), want[c]), (mode, layout, c)
            g = cls(Q15, fo, R, M, N, n_channels=C, layout=layout)
            g.set_state(blob)
            b2 = g.run(xin[:, half:] if layout == "planar" else xin[half:])
            assert np.array_equal(b2, b)


@pytest.mark.parametrize("recursive", ["0", "1"])
def test_cic_device_path(engine, oracle, recursive, monkeypatch):
    """BASELINE config 3 on the fast decimator, both evaluation forms: the non-recursive three-stage half-band cascade
    (DP2A first stage) and the integrator / comb recursion with run-in; interleaved IQ and planar, ragged calls."""
    import torch
    monkeypatch.setenv("B2D_CIC_RECURSIVE", recursive)
    rng = np.random.default_rng(22)
    x = rng.integers(-32768, 32767, size=(1 << 20, 2), endpoint=True).astype(np.int16)
    want = [oracle.CicB("dec", Q15, (28, 13), 8, 1, 4).run(x[:, c]) for c in range(2)]
    f = engine.ac_cic_dec_full(Q15, (28, 13), 8, 1, 4, n_channels=2, layout="interleaved")
    assert f.path == "cic_fast"
    y = f.run(torch.from_numpy(x).cuda()).cpu().numpy()
    for c in range(2):
        assert np.array_equal(y[c].astype(np.int64), want[c])
    f.reset()
    cuts = [0, 8, 4096 + 8, 300000, 300008, 1 << 20]           # multiples of R keep the calls on the fast kernel
    xd = torch.from_numpy(x).cuda()
    parts = [f.run(xd[a:b]).cpu().numpy() for a, b in zip(cuts[:-1], cuts[1:])]
    for c in range(2):
        assert np.array_equal(np.concatenate([p[c] for p in parts]).astype(np.int64), want[c])
    g = engine.ac_cic_dec_full(Q15, (28, 13), 8, 1, 4, n_channels=2, layout="planar")
    yp = g.run(torch.from_numpy(np.ascontiguousarray(x.T)).cuda()).cpu().numpy()
    for c in range(2):
        assert np.array_equal(yp[c].astype(np.int64), want[c])
    for kind in ("min", "max", "alt"):
        xe = oracle.rand_raw(rng, Q15, 40000, kind).astype(np.int16)
        h = engine.ac_cic_dec_full(Q15, (28, 13), 8, 1, 4)
        assert np.array_equal(h.run(torch.from_numpy(xe).cuda()).cpu().numpy().astype(np.int64),
                              oracle.CicB("dec", Q15, (28, 13), 8, 1, 4).run(xe)), kind


@pytest.mark.parametrize("staged", ["0", "1"])
@pytest.mark.parametrize("N,M,fo", [(3, 1, (20, 5)), (4, 1, (22, 7)), (3, 2, (23, 8))])
def test_cic_intr_fast_path_device(engine, oracle, staged, N, M, fo, monkeypatch):
    """BASELINE config 5 first stage (R = 4) on the polyphase kernels -- the shuffle-window kernel (all four output
    alignments A = out_first mod 4 occur across the calls below) and the shared-memory staged one: device path, call
    boundaries inside an input period, chunks smaller than a warp window, DC gain."""
    import torch
    monkeypatch.setenv("B2D_CIC_INTR_STAGED", staged)
    rng = np.random.default_rng(23 + N + M)
    n = 300001
    x = rng.integers(-32768, 32767, size=n, endpoint=True).astype(np.int16)
    f = engine.ac_cic_intr_full(Q15, fo, 4, M, N)
    assert f.path == "cic_intr_fast"
    xd = torch.from_numpy(x).cuda()
    cuts = [0, 1, 2, 3, 5, 40, 100003, 100004, 100010, 250000, n]
    parts = [f.run(xd[a:b]).cpu().numpy() for a, b in zip(cuts[:-1], cuts[1:])]
    want = oracle.CicB("intr", Q15, fo, 4, M, N).run(x)
    got = np.concatenate(parts).astype(np.int64)
    assert got.shape == want.shape and np.array_equal(got, want)
    for kind in ("min", "max", "alt"):          # extremes: the lossless width is exactly filled
        f.reset()
        xe = oracle.rand_raw(rng, Q15, 4097, kind).astype(np.int16)
        assert np.array_equal(f.run(torch.from_numpy(xe).cuda()).cpu().numpy().astype(np.int64),
                              oracle.CicB("intr", Q15, fo, 4, M, N).run(xe)), kind
    f.reset()
    ydc = f.run(torch.full((5000,), 7, dtype=torch.int16, device="cuda"))
    assert int(ydc[-1]) == 7 * (4 * M) ** N // 4          # DC gain of the interpolator: (R*M)^N / R


# ------------------------------------------------------------------------ full-size properties
def test_full_size_fir_windows(engine, oracle, ovs_auto):
    """BASELINE config 2 at full size (2^30 IQ samples, 256 taps): an output depends on a 256-sample window only,
    so random windows of the device result are re-derived by the oracle from the same inputs."""
    import torch
    n = 1 << 30
    g = torch.Generator(device="cuda").manual_seed(20260101)
    x = torch.randint(-32768, 32768, (n, 2), dtype=torch.int16, device="cuda", generator=g)
    rng = np.random.default_rng(20260101)
    h = oracle.rand_raw(rng, Q15, 256)
    f = make_fir(engine, "load", Q15, Q15, ACC40, ACC40, 256, "SHIFT_REG", h, n_channels=2, layout="interleaved")
    assert f.path == ("fir_ovs" if ovs_auto == "auto" else "fir_q15")
    y = f.run(x)
    torch.cuda.synchronize()
    assert y.shape == (n, 2)
    starts = [0, 1, 4096 - 300, 3840 - 100, n - 2000] + [int(v) for v in rng.integers(300, n - 3000, size=24)]
    for s in starts:
        lo = max(0, s - 255)
        xs = x[lo:s + 1500].cpu().numpy()
        ys = y[s:s + 1500].cpu().numpy()
        for c in range(2):
            want = oracle_fir(oracle, Q15, Q15, ACC40, ACC40, 256, "SHIFT_REG", h, xs[:, c])
            assert np.array_equal(ys[:, c], want[s - lo:]), (s, c)
    del y
    # linearity in the coefficients at full size: filter(h1 + h2) == filter(h1) + filter(h2) (mod 2^40), checksummed
    h1 = oracle.rand_raw(rng, (15, 1), 256)
    h2 = oracle.rand_raw(rng, (15, 1), 256)
    sums = []
    for hh in (h1, h2, h1 + h2):
        f.run(None, hh, True)
        f.reset()
        sums.append(int(f.run(x).sum().item()))
    assert (sums[0] + sums[1] - sums[2]) % (1 << 40) == 0


def test_full_size_cic_windows(engine, oracle):
    """BASELINE config 3 at full size (2^30 IQ inputs, R=8, N=4): windows re-derived by the oracle + DC-gain checksum."""
    import torch
    n = 1 << 30
    g = torch.Generator(device="cuda").manual_seed(20260102)
    x = torch.randint(-32768, 32768, (n, 2), dtype=torch.int16, device="cuda", generator=g)
    f = engine.ac_cic_dec_full(Q15, (28, 13), 8, 1, 4, n_channels=2, layout="interleaved")
    y = f.run(x)
    torch.cuda.synchronize()
    assert y.shape == (2, n // 8)
    rng = np.random.default_rng(5)
    for m in [0, 1, 5, n // 8 - 700] + [int(v) for v in rng.integers(10, n // 8 - 1000, size=24)]:
        m0 = max(0, m - 8)                      # N*M low-rate samples of run-in make the restarted oracle exact
        xs = x[m0 * 8:(m + 600) * 8].cpu().numpy()
        for c in range(2):
            want = oracle.CicB("dec", Q15, (28, 13), 8, 1, 4).run(xs[:, c])
            assert np.array_equal(y[c, m:m + 600].cpu().numpy().astype(np.int64), want[m - m0:]), (m, c)
    # sum of all outputs == boxcar^4 gain-weighted input sum (mod 2^28) is awkward at the stream end; use DC instead
    f.reset()
    ydc = f.run(torch.full((1 << 24, 2), 3, dtype=torch.int16, device="cuda"))
    assert int(ydc[0, -1]) == 3 * 8 ** 4 and int(ydc[1, 1000]) == 3 * 8 ** 4


def test_full_size_cascade_and_prog1024_windows(engine, oracle, ovs_auto):
    """BASELINE configs[4] (interpolator R=4 N=3 + 63-tap FIR, one channel of 2^26 inputs = one GPU's share) and
    configs[3] (1024 taps, one GPU's share: 8 channels x 2^27) at full size: random output windows re-derived by the
    oracle from the same inputs, plus output count and a zero-input / DC property."""
    import torch
    rng = np.random.default_rng(20260104)
    # ---- configs[4]
    n = 1 << 26
    g = torch.Generator(device="cuda").manual_seed(20260104)
    x = torch.randint(-32768, 32768, (n,), dtype=torch.int16, device="cuda", generator=g)
    hc = oracle.rand_raw(rng, Q15, 63)
    f = engine.cic_intr_fir_cascade(Q15, (20, 5), 4, 1, 3, ACC40, Q15, ACC40, 63, "SHIFT_REG", coeffs=hc)
    y = f.run(x)
    torch.cuda.synchronize()
    assert f.path == "cicfir_fused" and y.shape == ((n - 1) * 4 + 1 - 2,)        # (K-1)R + 1 - (N-1): SURVEY.md 8a a15
    for k in [0, 3, n - 600] + [int(v) for v in rng.integers(100, n - 1000, size=16)]:
        k0 = max(0, k - 40)                         # 18 composite taps per phase + the comb run-in
        xs = x[k0:k + 500].cpu().numpy()
        oc, of = oracle.CicB("intr", Q15, (20, 5), 4, 1, 3), oracle.FirB((20, 5), Q15, ACC40, ACC40, 63, "SHIFT_REG")
        of.load(hc)
        want = of.run(oc.run(xs))                   # restarted k0 inputs early: exact once 63 + run-in outputs have passed
        o_lo = (k - k0) * 4 + 100                   # compare from 100 outputs into period k
        ref = want[o_lo:o_lo + 1500] if k0 > 0 else want[:1500]
        first = (k * 4 + 100) if k0 > 0 else 0      # both streams drop their first N-1 outputs: index i of the restart is 4*k0 + i here
        got = y[first:first + ref.size].cpu().numpy().astype(np.int64)
        assert np.array_equal(got, ref), k
    del y, f
    # ---- configs[3], per-GPU share
    C, n = 8, 1 << 27
    x = torch.randint(-32768, 32768, (C, n), dtype=torch.int16, device="cuda", generator=g)
    h = oracle.rand_raw(rng, Q15, 1024)
    f = make_fir(engine, "prog", Q15, Q15, ACC40, ACC40, 1024, "SHIFT_REG", h, n_channels=C, layout="planar")
    y = f.run(x)
    torch.cuda.synchronize()
    assert f.path == ("fir_ovs" if ovs_auto == "auto" and f.ovs_margin()[0] < 0.49 else "fir_q15") and y.shape == (C, n)
    for s in [0, 3072 - 50, n - 1500] + [int(v) for v in rng.integers(1100, n - 3000, size=10)]:
        lo = max(0, s - 1023)
        c = int(rng.integers(0, C))
        want = oracle_fir(oracle, Q15, Q15, ACC40, ACC40, 1024, "SHIFT_REG", h, x[c, lo:s + 800].cpu().numpy())
        assert np.array_equal(y[c, s:s + 800].cpu().numpy(), want[s - lo:]), (s, c)


def test_full_size_polydec_polyintr_intgdump_windows(engine, oracle):
    """Rows N2 / N4 at the bench sizes: windows re-derived by the oracle."""
    import torch
    rng = np.random.default_rng(20260105)
    g = torch.Generator(device="cuda").manual_seed(20260105)
    # polyphase decimator, 32 x 8 taps, 2^28 IQ inputs
    n = 1 << 28
    x = torch.randint(-32768, 32768, (n, 2), dtype=torch.int16, device="cuda", generator=g)
    hp = oracle.rand_raw(rng, Q15, 256)
    f = engine.ac_poly_dec(Q15, Q15, ACC40, ACC40, 32, 8, coeffs=hp, n_channels=2, layout="interleaved")
    y = f.run(x)
    assert f.path == "polydec_q15" and y.shape == (2, n // 8)
    for m in [0, n // 8 - 300] + [int(v) for v in rng.integers(100, n // 8 - 1000, size=12)]:
        m0 = max(0, m - 32)
        xs = x[m0 * 8:(m + 250) * 8].cpu().numpy()
        for c in range(2):
            ob = oracle.PdB(Q15, Q15, ACC40, ACC40, 32, 8)
            ob.load(hp)
            assert np.array_equal(y[c, m:m + 250].cpu().numpy().astype(np.int64), ob.run(xs[:, c])[m - m0:]), (m, c)
    del x, y, f
    # polyphase interpolator, 16 taps x 4 phases, 2^26 inputs
    n = 1 << 26
    x = torch.randint(-32768, 32768, (n,), dtype=torch.int16, device="cuda", generator=g)
    hi = oracle.rand_raw(rng, Q15, 64)
    f = engine.ac_poly_intr(Q15, Q15, ACC40, ACC40, 16, 4, "FOLD_ANTI", coeffs=hi)
    y = f.run(x)
    assert f.path == "polyintr_q15" and y.shape == (4 * n,)
    for k in [0, n - 400] + [int(v) for v in rng.integers(100, n - 1000, size=12)]:
        k0 = max(0, k - 16)
        ob = oracle.PiB(Q15, Q15, ACC40, ACC40, 16, 4, "FOLD_ANTI")
        ob.load(hi)
        want = ob.run(x[k0:k + 300].cpu().numpy())
        assert np.array_equal(y[4 * k:4 * (k + 300)].cpu().numpy().astype(np.int64), want[4 * (k - k0):]), k
    del y, f
    # integrate-and-dump, 4 channels x 256 samples per dump, 2^28 samples: every output is a plain sum
    f = engine.ac_intg_dump(Q15, (32, 17), (32, 17), 1024, 4)
    xs = x[: 1 << 26]
    frames = xs.numel() // 1024
    yd = f.run(xs, np.full(frames, 256))
    assert f.path == "intgdump_vec" and yd.shape == (frames, 4)
    assert torch.equal(yd, xs.view(frames, 256, 4).to(torch.int32).sum(dim=1).to(torch.int32))


def test_comm_single_rank_broadcast(engine, oracle):
    """The NCCL coefficient broadcast with a world of one (the multi-rank path is exercised by bench.py --gpus N)."""
    uid = engine.Comm.unique_id()
    comm = engine.Comm(uid, 0, 1)
    comm.barrier()
    rng = np.random.default_rng(31)
    h = oracle.rand_raw(rng, Q15, 64)
    x = oracle.rand_raw(rng, Q15, 5000)
    f = engine.ac_fir_load_coeffs(Q15, ACC40, Q15, ACC40, 64, "SHIFT_REG", comm=comm, root=0)
    f.load(h)
    assert np.array_equal(f.run(x.astype(np.int16)).astype(np.int64), oracle_fir(oracle, Q15, Q15, ACC40, ACC40, 64, "SHIFT_REG", h, x))
    f.close()
    comm.close()


# ------------------------------------------------------------------------ cascade: CIC interpolator -> FIR (config 5)
def oracle_cascade(oracle, fin, fmid, R, M, N, fc, fa, fo, taps, ft, h, chunks):
    cic = oracle.CicB("intr", fin, fmid, R, M, N)
    fir = oracle.FirB(fmid, fc, fa, fo, taps, ft)
    fir.load(h)
    return np.concatenate([fir.run(cic.run(x)) for x in chunks])


CASCADES = [  # (R, M, N, MID, taps, ftype, coeff fmt, expected path)
    (4, 1, 3, (20, 5), 63, "SHIFT_REG", Q15, "cicfir_fused"),          # BASELINE config 5
    (4, 1, 3, (20, 5), 63, "FOLD_ODD", Q15, "cicfir_fused"),
    (4, 1, 3, (20, 5), 16, "FOLD_EVEN", Q15, "cicfir_fused"),
    (4, 1, 3, (20, 5), 1, "SHIFT_REG", Q15, "cicfir_fused"),
    (4, 1, 3, (24, 9), 30, "TRANSPOSED", (12, 1), "cicfir_two_stage"), # loadable TRANSPOSED keeps partial sums across a change: FIR object
    (4, 1, 3, (24, 9), 30, "C_BUFF", (12, 1), "cicfir_fused"),         # wider MID than the lossless type
    (4, 1, 3, (20, 5), 21, "SHIFT_REG", (8, 1), "cicfir_fused"),          # composite taps fit 16 bits: two byte planes
    (2, 1, 4, (19, 4), 33, "C_BUFF", Q15, "cicfir_fused"),
    (2, 2, 3, (21, 6), 8, "SHIFT_REG", (10, 2, False), "cicfir_fused"),
    (8, 1, 2, (19, 4), 40, "ROTATE_SHIFT", Q15, "cicfir_fused"),
    (8, 1, 4, (25, 10), 63, "SHIFT_REG", Q15, "cicfir_two_stage"),     # composite taps exceed 24 bits
    (4, 1, 3, (18, 3), 63, "SHIFT_REG", Q15, "cicfir_two_stage"),      # MID narrower than lossless: the CIC output wraps
    (3, 1, 3, (20, 5), 20, "SHIFT_REG", Q15, "cicfir_two_stage"),      # R not in {2, 4, 8}
]


@pytest.mark.parametrize("case", CASCADES, ids=lambda c: f"R{c[0]}M{c[1]}N{c[2]}-{c[4]}{c[5]}-{c[7]}")
@pytest.mark.parametrize("two_stage", ["0", "1"])
def test_cic_fir_cascade(engine, oracle, case, two_stage, monkeypatch):
    """One object for ac_cic_intr_full -> ac_fir_*: identical to the two oracle objects in sequence, whole stream and
    chunked (call boundaries inside an input period), through the fused polyphase kernel and through the two-stage path."""
    R, M, N, mid, taps, ft, fc, want_path = case
    monkeypatch.setenv("B2D_CICFIR_TWO_STAGE", two_stage)
    rng = np.random.default_rng(R * 1000 + N * 100 + taps)
    n = 20011
    x = oracle.rand_raw(rng, Q15, n)
    h = oracle.rand_raw(rng, fc, taps)
    if ft in ("FOLD_EVEN", "FOLD_ODD"):
        h = np.concatenate([h[: (taps + 1) // 2], h[: taps // 2][::-1]])
    f = engine.cic_intr_fir_cascade(Q15, mid, R, M, N, ACC40, fc, ACC40, taps, ft, coeffs=h)
    assert f.path == (want_path if two_stage == "0" else "cicfir_two_stage")
    want = oracle_cascade(oracle, Q15, mid, R, M, N, fc, ACC40, ACC40, taps, ft, h, [x])
    y = f.run(x)
    assert y.shape == want.shape and np.array_equal(y.astype(np.int64), want), f.path
    f.reset()
    cuts = [0, 1, 2, 3, 7, 8, 1000, 1001, 9999, n]
    parts = [f.run(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate(parts).astype(np.int64), want)
    if ft == "TRANSPOSED" and two_stage == "0":
        # the constant-coefficient class can never change its taps: fused; the loadable one follows the reference's
        # partial sums across a change mid-stream (ac_fir_load_coeffs.h:265-278)
        g = engine.cic_intr_fir_cascade(Q15, mid, R, M, N, ACC40, fc, ACC40, taps, ft, coeffs=h, fir_class="const")
        assert g.path == "cicfir_fused"
        assert np.array_equal(g.run(x).astype(np.int64), want)
        h2 = oracle.rand_raw(rng, fc, taps)
        oc, of = oracle.CicB("intr", Q15, mid, R, M, N), oracle.FirB(mid, fc, ACC40, ACC40, taps, ft)
        of.load(h)
        w = [of.run(oc.run(x[:5000]))]
        of.load(h2)
        w.append(of.run(oc.run(x[5000:])))
        f.reset()
        f.load(h)
        y2 = [f.run(x[:5000])]
        f.load(h2)
        y2.append(f.run(x[5000:]))
        assert np.array_equal(np.concatenate(y2).astype(np.int64), np.concatenate(w))


def test_cic_fir_cascade_channels_device_and_extremes(engine, oracle):
    import torch
    rng = np.random.default_rng(99)
    C, n = 3, 70001
    h = oracle.rand_raw(rng, Q15, 63)
    x = rng.integers(-32768, 32767, size=(C, n), endpoint=True).astype(np.int16)
    want = [oracle_cascade(oracle, Q15, (20, 5), 4, 1, 3, Q15, ACC40, ACC40, 63, "SHIFT_REG", h, [x[c]]) for c in range(C)]
    for layout in ("planar", "interleaved"):
        f = engine.cic_intr_fir_cascade(Q15, (20, 5), 4, 1, 3, ACC40, Q15, ACC40, 63, "SHIFT_REG", coeffs=h, n_channels=C, layout=layout)
        assert f.path == "cicfir_fused"
        xin = torch.from_numpy(x if layout == "planar" else np.ascontiguousarray(x.T)).cuda()
        half = 33333
        a = f.run(xin[:, :half] if layout == "planar" else xin[:half]).cpu().numpy()
        b = f.run(xin[:, half:] if layout == "planar" else xin[half:]).cpu().numpy()
        for c in range(C):
            assert np.array_equal(np.concatenate([a[c], b[c]]), want[c]), (layout, c)
    # extremes: all-min samples x all-min taps exercise the accumulator wrap and every byte plane's sign handling
    for kx, kh in (("min", "min"), ("max", "min"), ("alt", "max")):
        xe, he = oracle.rand_raw(rng, Q15, 5000, kx), oracle.rand_raw(rng, Q15, 63, kh)
        f = engine.cic_intr_fir_cascade(Q15, (20, 5), 4, 1, 3, ACC40, Q15, ACC40, 63, "SHIFT_REG", coeffs=he)
        assert np.array_equal(f.run(xe).astype(np.int64),
                              oracle_cascade(oracle, Q15, (20, 5), 4, 1, 3, Q15, ACC40, ACC40, 63, "SHIFT_REG", he, [xe])), (kx, kh)
    with pytest.raises(engine.B2dError):
        engine.cic_intr_fir_cascade(Q15, (20, 5), 4, 1, 3, ACC40, Q15, ACC40, 63, coeffs=None).run(np.zeros(8, dtype=np.int16))


# ------------------------------------------------------------------------ ac_fir_reg_share
@pytest.mark.parametrize("cid", range(len(rc.RS_CONFIGS)), ids=lambda i: f"rs{i}-{rc.RS_CONFIGS[i][8]}-{rc.RS_CONFIGS[i][0]}")
def test_reg_share_vs_reference_outputs(engine, cid, path):
    """The engine against the committed outputs of the UNMODIFIED reference class ac_fir_reg_share: block run(),
    ac_firProgCoeffs_delay_line, and the scalar explicit-delay-line form the facade uses."""
    g = golden("rs_outputs.npz")
    N, fi, fo, fc, fa, mww, bs, bo, ft = rc.RS_CONFIGS[cid]
    x, ram, want = g[f"rs{cid}_x"], g[f"rs{cid}_ram"], g[f"rs{cid}_y"]
    f = engine.ac_fir_reg_share(N, fi, fo, fc, fa, mww, bs, bo, ft)
    y = np.concatenate([f.run(x[:7], ram), f.run(x[7:8]), f.run(x[8:], ram)])
    assert np.array_equal(y.astype(np.int64), want), f.path
    assert int(f.delay_line()) == int(g[f"rs{cid}_dl"][0])
    # scalar form: the caller owns the delay line (reg[0] newest); three positions inside the stream
    for n in (0, 5, x.size - 1):
        reg = np.zeros(N, dtype=np.int64)
        hist = x[max(0, n - N + 1): n + 1][::-1]
        reg[: hist.size] = hist
        assert int(f.run_window(reg)) == int(want[n]), n


def test_reg_share_random_and_errors(engine, oracle, ovs):
    rng = np.random.default_rng(41)
    for N, ft, fc in ((256, "FOLD_EVEN_ANTI", (15, 1)), (255, "FOLD_ODD_ANTI", (15, 1)), (256, "FOLD_EVEN_ANTI", Q15), (64, "SHIFT_REG", Q15)):
        f = engine.ac_fir_reg_share(N, Q15, ACC40, fc, ACC40, 1, 1, 0, ft, n_channels=2, layout="interleaved")
        assert f.path == ("fir_q15" if (fc != Q15 or ft == "SHIFT_REG") else "fir_wide"), f.path   # negated 16-bit taps need 17 bits
        ob = [oracle.RsB(Q15, ACC40, fc, ACC40, N, 1, 1, 0, ft) for _ in range(2)]
        ram = oracle.rand_raw(rng, fc, ob[0].ram_words)
        x = rng.integers(-32768, 32767, size=(9000, 2), endpoint=True).astype(np.int16)
        y = f.run(x, ram)
        for c in range(2):
            assert np.array_equal(y[:, c].astype(np.int64), ob[c].run(x[:, c], ram)), (N, ft, c)
        dl = f.delay_line()
        assert [int(v) for v in dl] == [ob[c].delay_out() for c in range(2)]
        if f.path != "fir_wide":
            check_ovs(f, ovs, N)                  # the anti-symmetric folds through the overlap-save evaluation as well
    with pytest.raises(engine.B2dError):      # ac_fir_reg_share does not dispatch TRANSPOSED
        engine.ac_fir_reg_share(16, Q15, ACC40, Q15, ACC40, 1, 1, 0, "TRANSPOSED")
    with pytest.raises(engine.B2dError):      # ... and the const / load / prog classes do not dispatch _ANTI
        engine.ac_fir_load_coeffs(Q15, ACC40, Q15, ACC40, 16, "FOLD_EVEN_ANTI")
    f = engine.ac_fir_reg_share(16, Q15, ACC40, Q15, ACC40, 4, 3, 0, "SHIFT_REG")
    with pytest.raises(engine.B2dError):      # 16 taps are not a whole number of 3-tap blocks
        f.load(np.zeros(64))
    f = engine.ac_fir_reg_share(16, Q15, ACC40, Q15, ACC40, 4, 2, 1, "SHIFT_REG")
    with pytest.raises(engine.B2dError):      # RAM image too small
        f.load(np.zeros(8))


# ------------------------------------------------------------------------ ac_poly_dec
@pytest.mark.parametrize("cid", range(len(rc.PD_CONFIGS)), ids=lambda i: f"pd{i}-NT{rc.PD_CONFIGS[i][4]}-DF{rc.PD_CONFIGS[i][5]}")
def test_poly_dec_vs_reference_outputs(engine, cid, path):
    """The engine against the committed outputs of the UNMODIFIED reference class ac_poly_dec (ragged calls: 1 sample,
    DF+1 samples, the rest -- incomplete groups stay pending between calls)."""
    g = golden("rs_outputs.npz")
    fi, fc, fa, fo, nt, df = rc.PD_CONFIGS[cid]
    x = g[f"pd{cid}_x"]
    f = engine.ac_poly_dec(fi, fc, fa, fo, nt, df, coeffs=g[f"pd{cid}_c"])
    y = np.concatenate([f.run(x[:1]), f.run(x[1:df + 2]), f.run(x[df + 2:])])
    assert np.array_equal(y.astype(np.int64), g[f"pd{cid}_y"]), f.path


def test_poly_dec_ddc_chain_channels_and_device(engine, oracle):
    """The R = 8 CIC decimator's partner: 32 taps x 8 phases = 256-tap decimate-by-8 on 16-bit IQ, device path, big tiles."""
    import torch
    rng = np.random.default_rng(55)
    nt, df, n = 32, 8, 400003
    c = oracle.rand_raw(rng, Q15, nt * df)
    x = rng.integers(-32768, 32767, size=(n, 2), endpoint=True).astype(np.int16)
    f = engine.ac_poly_dec(Q15, Q15, ACC40, ACC40, nt, df, coeffs=c, n_channels=2, layout="interleaved")
    assert f.path == "polydec_q15"
    xd = torch.from_numpy(x).cuda()
    cuts = [0, 3, 200000, 200001, n]
    parts = [f.run(xd[a:b]).cpu().numpy() for a, b in zip(cuts[:-1], cuts[1:])]
    for ch in range(2):
        ob = oracle.PdB(Q15, Q15, ACC40, ACC40, nt, df)
        ob.load(c)
        assert np.array_equal(np.concatenate([p[ch] for p in parts]), ob.run(x[:, ch])), ch
    # planar channels, odd tap count (padded phases), three accumulator flushes (5 phases x 112 padded taps), extremes
    nt2, df2 = 100, 5
    for kind in ("uniform", "min", "alt"):
        c2 = oracle.rand_raw(rng, Q15, nt2 * df2, "uniform" if kind == "alt" else kind)
        x2 = np.stack([oracle.rand_raw(rng, Q15, 30007, kind) for _ in range(3)]).astype(np.int16)
        f2 = engine.ac_poly_dec(Q15, Q15, ACC40, ACC40, nt2, df2, coeffs=c2, n_channels=3)
        assert f2.path == "polydec_q15"
        y2 = f2.run(x2)
        for ch in range(3):
            ob = oracle.PdB(Q15, Q15, ACC40, ACC40, nt2, df2)
            ob.load(c2)
            assert np.array_equal(y2[ch], ob.run(x2[ch])), (kind, ch)
    with pytest.raises(engine.B2dError):
        engine.ac_poly_dec(Q15, Q15, ACC40, ACC40, nt, df).run(np.zeros(16, dtype=np.int16))     # no coefficients yet
    with pytest.raises(engine.B2dError):
        engine.ac_poly_dec(Q15, Q15, ACC40, ACC40, nt, df, coeffs=np.zeros(nt))                  # needs NTAPS * DF values


# ------------------------------------------------------------------------ ac_intg_dump
@pytest.mark.parametrize("cid", range(len(rc.ID_CONFIGS)), ids=lambda i: f"id{i}-NS{rc.ID_CONFIGS[i][3]}-CHN{rc.ID_CONFIGS[i][4]}")
def test_intg_dump_vs_reference_outputs(engine, cid, path):
    """The engine against the committed outputs of the UNMODIFIED reference class ac_intg_dump: three calls with
    non-dumping frames in the middle and at the end, so the running sums carry across frames and across calls."""
    g = golden("rs_outputs.npz")
    fi, fa, fo, NS, CHN = rc.ID_CONFIGS[cid]
    x, ns, xlen = g[f"id{cid}_x"], g[f"id{cid}_ns"], g[f"id{cid}_xlen"]
    f = engine.ac_intg_dump(fi, fa, fo, NS, CHN)
    ys, o = [], 0
    for call in range(3):
        ys.append(f.run(x[o:o + xlen[call]], ns[6 * call:6 * call + 6]).reshape(-1))
        o += xlen[call]
    assert np.array_equal(np.concatenate(ys).astype(np.int64), g[f"id{cid}_y"])


# ------------------------------------------------------------------------ ac_mv_avg (parity unpinned)
@pytest.mark.parametrize("cid", range(len(rc.MV_CONFIGS)), ids=[f"mv{i}-{c[2]}-{c[1]}" for i, c in enumerate(rc.MV_CONFIGS)])
def test_mv_avg_vs_reference_outputs(engine, oracle, cid):
    """The engine against the committed outputs of the UNMODIFIED ac_mv_avg.h (driven over the restated window class):
    several bursts per call, the shortest legal burst, every accumulator / output mode of the table; then a large
    device-resident call against the restatement."""
    import torch
    g = golden("rs_outputs.npz")
    maxs, taps, wt, fi, fo, fa, fc = rc.MV_CONFIGS[cid]
    ns1, ns2 = (int(v) for v in g[f"mv{cid}_ns"])
    f = engine.ac_mv_avg(maxs, taps, wt, fi, fo, fa, fc, g[f"mv{cid}_c"])
    assert f.path == "mvavg_generic"
    assert np.array_equal(f.run(g[f"mv{cid}_x1"], ns1).astype(np.int64), g[f"mv{cid}_y1"])
    assert np.array_equal(f.run(g[f"mv{cid}_x2"], ns2).astype(np.int64), g[f"mv{cid}_y2"])      # nothing carries over
    rng = np.random.default_rng(100 + cid)
    x = oracle.rand_raw(rng, fi, 257 * maxs)
    xin = torch.from_numpy(x.astype(f._in_dt)).cuda()
    y = f.run(xin, maxs).cpu().numpy().astype(np.int64)
    assert np.array_equal(y, oracle.mv_run_b(fi, fo, fa, fc, taps, wt, g[f"mv{cid}_c"], x, maxs))
    with pytest.raises(engine.B2dError):
        f.run(x[: maxs + 1], maxs)                       # not a whole number of bursts
    with pytest.raises(engine.B2dError):
        f.run(x[: 2 * (maxs + 1)], maxs + 1)             # burst longer than MAX_SAMPLE
    if taps > 1:
        with pytest.raises(engine.B2dError):
            f.run(x[: taps - 1], taps - 1)               # window span above the burst length


def test_intg_dump_regular_frames_device(engine, oracle):
    """Constant frame length (the streaming case): 4 channels x 64 samples per dump on the device path, and the
    argument check that stands in for the reference reading past the end of its channel."""
    import torch
    rng = np.random.default_rng(66)
    NS, CHN, n, frames = 1024, 4, 64, 20000
    x = rng.integers(-32768, 32767, size=frames * n * CHN, endpoint=True).astype(np.int16)
    f = engine.ac_intg_dump(Q15, (32, 17), (32, 17), NS, CHN)
    y = f.run(torch.from_numpy(x).cuda(), np.full(frames, n)).cpu().numpy()
    want = oracle.IdB(Q15, (32, 17), (32, 17), NS, CHN).run(x, np.full(frames, n)).reshape(frames, CHN)
    assert np.array_equal(y.astype(np.int64), want)
    assert np.array_equal(want, x.reshape(frames, n, CHN).astype(np.int64).sum(axis=1))     # F_acc == F_in: plain sums
    with pytest.raises(engine.B2dError):
        f.run(x[:-1], np.full(frames, n))


@pytest.mark.parametrize("CHN,n,fi,fa", [
    (1, 8, Q15, (32, 17)), (1, 256, Q15, (32, 17)), (1, 1000, Q15, (32, 17)), (2, 4, Q15, (40, 25)), (2, 64, Q15, (40, 25)),
    (2, 324, Q15, (32, 17)), (4, 8, Q15, (32, 17)), (4, 64, Q15, (32, 17)), (4, 1024, Q15, (18, 3)), (8, 8, Q15, (18, 3)),
    (8, 32, Q15, (32, 17)), (8, 77, Q15, (32, 17)), (4, 64, Q15, (24, 12)), (2, 128, (16, 1, True), (24, 12, True, "AC_RND")),
    (4, 16, (12, 0, False), (20, 8, False)), (1, 2048, (16, 4, False), (30, 18, False)), (4, 2, Q15, (32, 17)), (8, 1, Q15, (32, 17)),
    (1, 24, Q15, (32, 17)), (4, 64 * 1024, Q15, (40, 25)),
], ids=lambda v: str(v).replace(" ", ""))
def test_intg_dump_vector_path_geometries(engine, oracle, CHN, n, fi, fa):
    """Equal frames on the device path over the geometries of intgdump_vec_kernel (sub-warp groups, whole warps, long
    segments, per-term truncation / rounding, unsigned samples) and the ones that must fall back; two calls so the
    second starts from a cleared accumulator, then a ragged third call that leaves a carry behind."""
    import torch
    rng = np.random.default_rng(1000 * CHN + n)
    NS = max(n, 16)
    frames = max(3, min(3000, (1 << 21) // (n * CHN)))
    W, S = fi[0], (fi[2] if len(fi) > 2 else True)
    lo, hi = (-(1 << (W - 1)), (1 << (W - 1)) - 1) if S else (0, (1 << W) - 1)
    f = engine.ac_intg_dump(fi, fa, fa, NS, CHN)
    ob = oracle.IdB(fi, fa, fa, NS, CHN)
    for call in range(2):
        x = rng.integers(lo, hi, size=frames * n * CHN, endpoint=True).astype(np.int16 if S else np.uint16)
        y = f.run(torch.from_numpy(x.view(np.int16)).cuda(), np.full(frames, n)).cpu().numpy()
        assert np.array_equal(y.astype(np.int64), ob.run(x, np.full(frames, n)).reshape(frames, CHN)), (call, f.path)
    L = n * CHN // 8
    vec = (n * CHN) % 8 == 0 and (L > 32 or ((L & (L - 1)) == 0 and L >= CHN)) and (fi[0] - fi[1]) - (fa[0] - fa[1]) <= 15
    assert f.path == ("intgdump_vec" if vec else ("intgdump_warp" if 32 % CHN == 0 else "intgdump_thread")), f.path
    tok = np.array([n, NS + 5, n, 0], dtype=np.uint32)     # ragged: two frames that do not dump
    m = (2 * n + 2 * NS) * CHN
    x = rng.integers(lo, hi, size=m, endpoint=True).astype(np.int16 if S else np.uint16)
    y = f.run(torch.from_numpy(x.view(np.int16)).cuda(), tok).cpu().numpy()
    assert np.array_equal(y.astype(np.int64), ob.run(x, tok).reshape(-1, CHN))
    x = rng.integers(lo, hi, size=frames * n * CHN, endpoint=True).astype(np.int16 if S else np.uint16)
    y = f.run(torch.from_numpy(x.view(np.int16)).cuda(), np.full(frames, n)).cpu().numpy()     # picks the carry up
    assert np.array_equal(y.astype(np.int64), ob.run(x, np.full(frames, n)).reshape(frames, CHN))


# ------------------------------------------------------------------------ ac_poly_intr
@pytest.mark.parametrize("cid", range(len(rc.PI_CONFIGS)),
                         ids=lambda i: f"pi{i}-{rc.PI_CONFIGS[i][6]}-NT{rc.PI_CONFIGS[i][4]}-IF{rc.PI_CONFIGS[i][5]}")
def test_poly_intr_vs_reference_outputs(engine, cid, path):
    """The engine against the committed outputs of the UNMODIFIED reference class ac_poly_intr: three runs, a reload of the
    coefficient and control structures half way (the parked accumulators keep the values of the old set, the new
    sign / corr apply at once), one more run."""
    from test_oracle import pi_replay
    g = golden("rs_outputs.npz")
    fi, fc, fa, fo, nt, IF, ft = rc.PI_CONFIGS[cid]
    f = engine.ac_poly_intr(fi, fc, fa, fo, nt, IF, ft)
    assert np.array_equal(pi_replay(f, g, cid).astype(np.int64), g[f"pi{cid}_y"]), f.path
    if path == "auto":
        want = "polyintr_q15" if (ft == "FOLD_ANTI" and fi[0] <= 16 and fa[0] == 40 and IF in (2, 4, 8)) else None
        assert want is None or f.path == want, f.path


@pytest.mark.parametrize("ft,nt,IF", [("FOLD_ANTI", 16, 4), ("FOLD_ANTI", 3, 2), ("FOLD_ANTI", 33, 8), ("FOLD_ANTI", 9, 5),
                                      ("FOLD_EVEN", 12, 4), ("FOLD_ODD", 11, 3)])
def test_poly_intr_random_channels_and_device(engine, oracle, ft, nt, IF, path):
    """Random streams against Oracle B: host and device paths, planar and interleaved channels with per-channel
    coefficient / control sets, calls of ragged lengths (including single samples and an empty call)."""
    import torch
    rng = np.random.default_rng(nt * 100 + IF)
    C = 3
    csz = rc.pi_coeffsz((None, None, None, None, nt, IF, ft))
    cs = [rng.integers(-32768, 32767, size=csz, endpoint=True) for _ in range(C)]
    sg = [rng.integers(0, 2, size=IF) for _ in range(C)]
    cr = [rng.integers(0, IF, size=IF) for _ in range(C)]
    n = 30000
    x = rng.integers(-32768, 32767, size=(C, n), endpoint=True).astype(np.int16)
    x[:, 100:140] = -32768
    want = []
    for c in range(C):
        ob = oracle.PiB(Q15, Q15, ACC40, ACC40, nt, IF, ft)
        ob.load(cs[c], sg[c], cr[c])
        want.append(ob.run(x[c]))
    want = np.stack(want)
    for layout in ("planar", "interleaved"):
        for dev in (False, True):
            f = engine.ac_poly_intr(Q15, Q15, ACC40, ACC40, nt, IF, ft, n_channels=C, layout=layout)
            for c in range(C):
                f.load(cs[c], sg[c], cr[c], channel=c)
            ys, o = [], 0
            for ln in (1, 0, 1, 7, 4096, 12345, n):
                xi = x[:, o:min(o + ln, n)]
                if layout == "interleaved":
                    xi = np.ascontiguousarray(xi.T)
                y = f.run(torch.from_numpy(np.ascontiguousarray(xi)).cuda()).cpu().numpy() if dev else f.run(xi)
                ys.append(np.asarray(y).reshape(C, -1))
                o = min(o + ln, n)
            assert np.array_equal(np.concatenate(ys, axis=1).astype(np.int64), want), (layout, dev, f.path)
    with pytest.raises(engine.B2dError):
        engine.ac_poly_intr(Q15, Q15, ACC40, ACC40, nt, IF, ft).run(np.zeros(4, dtype=np.int16))     # nothing loaded yet
    with pytest.raises(engine.B2dError):
        engine.ac_poly_intr(Q15, Q15, ACC40, ACC40, nt, IF, ft, coeffs=np.zeros(csz), corr=np.full(IF, IF))   # corr out of range


# ------------------------------------------------------------------------ checkpoint / resume of the later handle types
@pytest.mark.parametrize("two_stage", ["0", "1"])
def test_checkpoint_resume_cascade_polydec_polyintr_intgdump(engine, oracle, two_stage, monkeypatch):
    """all filter state lives in object members; get_state after k samples + set_state into a fresh
    handle must continue the stream exactly like the uninterrupted object (and a blob of another configuration is
    refused)."""
    monkeypatch.setenv("B2D_CICFIR_TWO_STAGE", two_stage)
    rng = np.random.default_rng(99)
    x = rng.integers(-32768, 32767, size=5000, endpoint=True).astype(np.int16)
    cut = 1237
    # cascade (BASELINE configs[4])
    g = rng.integers(-32768, 32767, size=63, endpoint=True).astype(np.int16)
    mk = lambda: engine.cic_intr_fir_cascade(Q15, (20, 5), 4, 1, 3, ACC40, Q15, ACC40, 63, "SHIFT_REG", coeffs=g)
    a, b = mk(), mk()
    whole = a.run(x)
    first = b.run(x[:cut])
    c = mk()
    c.set_state(b.get_state())
    assert np.array_equal(np.concatenate([first, c.run(x[cut:])]), whole), a.path
    with pytest.raises(engine.B2dError):
        engine.cic_intr_fir_cascade(Q15, (20, 5), 4, 1, 3, ACC40, Q15, ACC40, 31, "SHIFT_REG", coeffs=g[:31]).set_state(b.get_state())
    if two_stage == "1":
        return
    # polyphase decimator
    hp = rng.integers(-32768, 32767, size=64, endpoint=True).astype(np.int16)
    mk = lambda: engine.ac_poly_dec(Q15, Q15, ACC40, ACC40, 16, 4, coeffs=hp)
    a, b, c = mk(), mk(), mk()
    whole, first = a.run(x), b.run(x[:cut])
    c.set_state(b.get_state())
    assert np.array_equal(np.concatenate([first, c.run(x[cut:])]), whole)
    # polyphase interpolator, folded form: the parked accumulators and `init` travel with the blob
    for ft, nt in (("FOLD_EVEN", 8), ("FOLD_ANTI", 16)):
        csz = rc.pi_coeffsz((None, None, None, None, nt, 4, ft))
        hc = rng.integers(-32768, 32767, size=csz, endpoint=True).astype(np.int16)
        mk = lambda: engine.ac_poly_intr(Q15, Q15, ACC40, ACC40, nt, 4, ft, coeffs=hc, sign=[1, 0, 0, 1], corr=[3, 1, 2, 0])
        a, b, c = mk(), mk(), mk()
        whole, first = a.run(x), b.run(x[:cut])
        c.set_state(b.get_state())
        assert np.array_equal(np.concatenate([first, c.run(x[cut:])]), whole), ft
        with pytest.raises(engine.B2dError):
            engine.ac_poly_intr(Q15, Q15, ACC40, ACC40, nt + 2, 4, ft).set_state(b.get_state())
    # integrate-and-dump: a frame that does not dump leaves running sums behind
    mk = lambda: engine.ac_intg_dump(Q15, (32, 17), (32, 17), 64, 4)
    a, b, c = mk(), mk(), mk()
    tok1, tok2 = np.array([64, 99, 0]), np.array([32, 64])
    n1, n2 = (64 + 64 + 64) * 4, (32 + 64) * 4
    whole = np.concatenate([a.run(x[:n1], tok1), a.run(x[n1:n1 + n2], tok2)])
    first = b.run(x[:n1], tok1)
    c.set_state(b.get_state())
    assert np.array_equal(np.concatenate([first, c.run(x[n1:n1 + n2], tok2)]), whole)


# ------------------------------------------------------------------------------------------------ packed host-link format
@pytest.mark.gpu
@pytest.mark.parametrize("chunk_bytes", ["", "40000"])
def test_packed_wire_format_matches_containers(engine, oracle, chunk_bytes, monkeypatch):
    """B2D_WIRE_PACKED (b200dsp.h): the host-buffer run() writes ceil(W_out / 8) little-endian bytes per value, same
    element order; widened back on the host it is the container result (and the oracle's), for every handle family,
    across several pipeline chunks and with counts that are not multiples of the pack kernel's group of four."""
    if chunk_bytes:
        monkeypatch.setenv("B2D_PIPE_CHUNK_BYTES", chunk_bytes)     # dozens of chunks through the three pipeline slots
    rng = np.random.default_rng(2026)
    n = 70001
    xi = oracle.rand_raw(rng, Q15, 2 * n).astype(np.int16).reshape(n, 2)
    h = oracle.rand_raw(rng, Q15, 64)
    f = engine.ac_fir_load_coeffs(Q15, ACC40, Q15, ACC40, 64, "SHIFT_REG", n_channels=2, layout="interleaved")
    f.load(h)
    want = f.run(xi)
    f.reset()
    f.set_wire("packed")
    yp = f.run(xi)
    assert yp.dtype == np.uint8 and yp.shape == (n, 2, 5)
    assert np.array_equal(f.unpack_wire(yp), want)
    ob = oracle.FirB(Q15, Q15, ACC40, ACC40, 64, "SHIFT_REG")
    ob.load(h)
    assert np.array_equal(want[:, 1].astype(np.int64), ob.run(xi[:, 1]))
    f.set_wire("container")
    f.reset()
    assert np.array_equal(f.run(xi), want)
    # planar multi-channel FIR with a 24-bit output (3 bytes on the wire, 4 in the container)
    xp = oracle.rand_raw(rng, Q15, 3 * 9999).astype(np.int16).reshape(3, 9999)
    g = engine.ac_fir_load_coeffs(Q15, (24, 4), Q15, (30, 6), 16, "C_BUFF", n_channels=3)
    g.load(h[:16])
    wg = g.run(xp)
    g.reset()
    g.set_wire("packed")
    assert np.array_equal(g.unpack_wire(g.run(xp)), wg)
    # rate-changing families: planar outputs with the call's output count as stride
    x1 = xi[:, 0].copy()
    cases = [
        (engine.ac_cic_dec_full(Q15, (20, 5), 8, 1, 4, n_channels=2, layout="interleaved"), xi),
        (engine.ac_cic_intr_full(Q15, (20, 5), 4, 1, 3), x1[:20001]),
        (engine.cic_intr_fir_cascade(Q15, (20, 5), 4, 1, 3, ACC40, Q15, ACC40, 63, "SHIFT_REG", coeffs=h[:63]), x1[:20001]),
        (engine.ac_poly_dec(Q15, Q15, ACC40, ACC40, 8, 8, coeffs=h), x1),
        (engine.ac_poly_intr(Q15, Q15, ACC40, ACC40, 16, 4, "FOLD_ANTI", coeffs=h), x1[:20001]),
    ]
    for obj, x in cases:
        monkeypatch.delenv("B2D_PIPE_CHUNK_BYTES", raising=False)
        w = obj.run(x)                                  # one chunk, containers
        if chunk_bytes:
            monkeypatch.setenv("B2D_PIPE_CHUNK_BYTES", chunk_bytes)
            obj.reset()
            assert np.array_equal(obj.run(x), w), type(obj).__name__      # many chunks, containers
        obj.reset()
        obj.set_wire("packed")
        got = obj.run(x)
        assert got.dtype == np.uint8 and got.shape[:-1] == w.shape, type(obj).__name__
        assert np.array_equal(obj.unpack_wire(got), w), type(obj).__name__


# ------------------------------------------------------------------------------------------------ fir_q24
Q24_CASES = [
    # (in, coeff, acc, out, taps, ftype, layout/channels): samples of 17..24 bits, taps <= 16 bits, exact-shift accumulators
    ((20, 5), Q15, ACC40, ACC40, 63, "SHIFT_REG", 1),                       # BASELINE configs[4], second stage
    ((20, 5), Q15, ACC40, ACC40, 63, "FOLD_ODD", 1),
    ((24, 4), Q15, (48, 12), (48, 12), 256, "C_BUFF", 1),                   # full 24-bit samples, two accumulation blocks
    ((24, 4), Q15, (48, 12), (20, 3, True, "AC_RND", "AC_SAT"), 300, "TRANSPOSED", 3),   # odd tap count, converted output, planar channels
    ((23, 3, False), (12, 2, False), (44, 10, False), (44, 10, False), 40, "FOLD_EVEN", 2),   # everything unsigned
    ((17, 1), (9, 1), ACC40, (24, 8), 1, "SHIFT_REG", 1),                   # one tap
    ((18, 2), (15, 1), (40, 9), (40, 9), 16, "ROTATE_SHIFT", 2),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", Q24_CASES, ids=lambda c: f"in{c[0][0]}-c{c[1][0]}-{c[4]}{c[5]}-C{c[6]}")
def test_fir_q24_path(engine, oracle, case):
    """Three sample byte planes x 16-bit coefficient lanes (fir_q24.cu) against the restatement: full-range random data,
    the extremes of every plane, chunked calls (history carry), multi-channel in both layouts, coefficient reload."""
    fi, fc, fa, fo, taps, ft, C = case
    rng = np.random.default_rng(taps * 7 + C)
    n = 9001
    h = oracle.rand_raw(rng, fc, taps)
    if ft in ("FOLD_EVEN", "FOLD_ODD"):
        h = np.concatenate([h[: (taps + 1) // 2], h[: taps // 2][::-1]])
    for kind in ("rand", "min", "max", "alt"):
        x = np.stack([oracle.rand_raw(rng, fi, n, kind) for _ in range(C)])
        want = []
        for c in range(C):
            ob = oracle.FirB(fi, fc, fa, fo, taps, ft)
            ob.load(h)
            want.append(ob.run(x[c]))
        want = np.stack(want)
        for layout in (("planar",) if C == 1 else ("planar", "interleaved")):
            f = engine.ac_fir_load_coeffs(fi, fo, fc, fa, taps, ft, n_channels=C, layout=layout)
            f.load(h)
            assert f.path == "fir_q24", f.path
            xin = x[0] if C == 1 else (x if layout == "planar" else np.ascontiguousarray(x.T))
            cuts = sorted({0, 1, 2, 9, taps + 3, 4096 + 5, n})
            parts = []
            for a, b in zip(cuts[:-1], cuts[1:]):
                seg = xin[a:b] if (C == 1 or layout == "interleaved") else xin[:, a:b]
                parts.append(np.atleast_1d(f.run(seg)))
            y = np.concatenate(parts, axis=(1 if (C > 1 and layout == "planar") else 0)).astype(np.int64)
            if C > 1 and layout == "interleaved":
                y = y.T
            assert np.array_equal(y.reshape(C, -1), want), (case, kind, layout)
    # the wide kernel and the generic kernel agree on the same configuration
    import os
    for force in ("2", "1"):
        os.environ["B2D_FORCE_GENERIC"] = force
        try:
            g = engine.ac_fir_load_coeffs(fi, fo, fc, fa, taps, ft)
            g.load(h)
            assert g.path in ("fir_wide", "fir_generic")
            ob = oracle.FirB(fi, fc, fa, fo, taps, ft)
            ob.load(h)
            xs = oracle.rand_raw(rng, fi, 700)
            assert np.array_equal(np.atleast_1d(g.run(xs)).astype(np.int64), ob.run(xs)), (case, force)
        finally:
            del os.environ["B2D_FORCE_GENERIC"]
