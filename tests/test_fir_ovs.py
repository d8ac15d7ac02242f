"""The overlap-save FIR path (ac_dsp_b200/csrc/fir_ovs.cu): long 16-bit filters through blocks of 4096 samples and an FP64
FFT, exact by an a-priori error bound evaluated on the loaded taps.

CPU: tests/cpp/fir_ovs_check.cu runs the kernel's thread phases as loops over thread ids and compares the rounded results
with a direct integer convolution (indexing, stream edges, the real-pair packing, the distance from an integer against
the bound).  GPU: the selection rules and the kernel itself against Oracle B (tests/test_gpu_parity.py runs its long q15
cases through both evaluations as well).
"""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q15, ACC40 = (16, 1), (40, 8)


def test_phases_on_the_cpu_against_direct_convolution(engine, tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    libdir = os.path.join(ROOT, "ac_dsp_b200", "lib")
    exe = str(tmp_path / "fir_ovs_check")
    subprocess.check_call([nvcc, "-std=c++17", "-O2", "-w", f"-I{ROOT}/ac_dsp_b200/csrc", os.path.join(ROOT, "tests", "cpp", "fir_ovs_check.cu"),
                           "-o", exe, f"-L{libdir}", "-lb200dsp", "-Xlinker", "-rpath", "-Xlinker", libdir])
    p = subprocess.run([exe], capture_output=True, text=True)
    print(p.stdout)
    assert p.returncode == 0 and "bad=0" in p.stdout, p.stdout[-3000:] + p.stderr[-1000:]
    # 500 blocks of +-full-scale samples against +-full-scale taps (random signs): the largest ||x||_2 ||h||_1 of the format
    p = subprocess.run([exe, "stress", "500"], capture_output=True, text=True)
    print(p.stdout)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-1000:]


def test_tables_and_spectrum_against_multiprecision(engine, tmp_path):
    """What the error bound assumes about the prepared constants: every twiddle component is the correctly rounded value
    (within 0.51 ulp) and every spectrum value is within 1.1 * 2^-53 * ||h||_1 / 4096 of the exact one."""
    mp = pytest.importorskip("mpmath")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    libdir = os.path.join(ROOT, "ac_dsp_b200", "lib")
    exe, blob = str(tmp_path / "fir_ovs_check"), str(tmp_path / "tables.bin")
    subprocess.check_call([nvcc, "-std=c++17", "-O2", "-w", f"-I{ROOT}/ac_dsp_b200/csrc", os.path.join(ROOT, "tests", "cpp", "fir_ovs_check.cu"),
                           "-o", exe, f"-L{libdir}", "-lb200dsp", "-Xlinker", "-rpath", "-Xlinker", libdir])
    subprocess.check_call([exe, "dump", blob])
    raw = open(blob, "rb").read()
    n = int(np.frombuffer(raw, dtype=np.int64, count=1)[0])
    h = np.frombuffer(raw, dtype=np.int64, count=n, offset=8)
    off = 8 + 8 * n
    tw1 = np.frombuffer(raw, dtype=np.float64, count=2 * 6 * 256, offset=off).reshape(6, 256, 2)
    off += tw1.nbytes
    tw2 = np.frombuffer(raw, dtype=np.float64, count=2 * 6 * 16, offset=off).reshape(6, 16, 2)
    off += tw2.nbytes
    hs = np.frombuffer(raw, dtype=np.float64, count=2 * 4096, offset=off).reshape(16, 256, 2)
    mp.mp.prec = 200
    u = mp.mpf(2) ** -53

    def ulp_err(got, exact):
        if exact == 0:
            return abs(mp.mpf(got))
        e = mp.floor(mp.log(abs(exact), 2))                 # 2^e <= |exact| < 2^(e+1): spacing of doubles there is 2^(e-52)
        return abs(mp.mpf(got) - exact) / mp.mpf(2) ** (e - 52)

    worst = mp.mpf(0)
    for i in range(6):
        mul = i + 1 if i < 3 else 4 * (i - 2)
        for tab, size, N in ((tw1, 256, 4096), (tw2, 16, 256)):
            for t in range(0, size, 1 if size == 16 else 5):
                ang = 2 * mp.pi * ((t * mul) % N) / N
                for got, exact in ((tab[i, t, 0], mp.cos(ang)), (tab[i, t, 1], -mp.sin(ang))):
                    if abs(exact) < mp.mpf(2) ** -60:                       # cos / sin of a multiple of pi / 2: an exact zero is expected
                        assert abs(got) < 1e-16
                    else:
                        worst = max(worst, ulp_err(got, exact))
    assert worst <= mp.mpf("0.51"), worst
    l1 = mp.mpf(int(np.abs(h).sum()))
    budget = mp.mpf("1.1") * u * l1 / 4096
    # position p = 16 c + j of the forward passes' output holds frequency (p >> 8) + 16 ((p >> 4) & 15) + 256 (p & 15); hs is stored [j][c]
    rng = np.random.default_rng(3)
    for p in [0, 1, 16, 255, 256, 4095] + [int(v) for v in rng.integers(0, 4096, size=40)]:
        f = (p >> 8) + 16 * ((p >> 4) & 15) + 256 * (p & 15)
        exact = mp.mpc(0)
        for k in range(n):
            ang = -2 * mp.pi * ((f * k) % 4096) / 4096
            exact += int(h[k]) * mp.mpc(mp.cos(ang), mp.sin(ang))
        exact /= 4096
        got = hs[p & 15, p >> 4]
        assert abs(mp.mpc(float(got[0]), float(got[1])) - exact) <= budget, (p, f)


def ofir(oracle, fi, fc, fa, fo, taps, ft, h, x):
    b = oracle.FirB(fi, fc, fa, fo, taps, ft)
    b.load(h)
    return b.run(x)


@pytest.fixture(params=["resid", "plain"])
def forced(request, monkeypatch):
    """Overlap-save for every call length; with the residual monitor (which takes the separate epilogue) and without (interior
    blocks run the last pass fused with the epilogue)."""
    monkeypatch.setenv("B2D_FIR_OVS", "2")
    if request.param == "resid":
        monkeypatch.setenv("B2D_OVS_RESID", "1")
    else:
        monkeypatch.delenv("B2D_OVS_RESID", raising=False)
    return request.param


def resid_ok(f, forced):
    bound, resid = f.ovs_margin()
    return bound < 0.49 and (0 <= resid < 0.01 if forced == "resid" else resid == -1.0)


@pytest.mark.gpu
def test_selection_rules(engine, oracle, monkeypatch):
    monkeypatch.delenv("B2D_FIR_OVS", raising=False)
    monkeypatch.setenv("B2D_OVS_RESID", "1")
    rng = np.random.default_rng(1)
    h = oracle.rand_raw(rng, Q15, 256)
    f = engine.ac_fir_load_coeffs(Q15, ACC40, Q15, ACC40, 256, "SHIFT_REG", n_channels=2, layout="interleaved")
    f.load(h, channel=0)
    f.load(h[::-1].copy(), channel=1)
    assert f.path == "fir_q15"                      # an IQ pair is one complex sequence: both channels need the same taps
    f.load(h, channel=1)
    assert f.path == "fir_ovs" and 0 < f.ovs_margin()[0] < 0.49
    # short calls still run the DP2A kernel, long ones overlap-save (1024 taps: from about 5 * 10^4 samples per call of an IQ
    # pair); the history crosses the switch in both directions
    h4 = oracle.rand_raw(rng, Q15, 1024)
    f4 = engine.ac_fir_load_coeffs(Q15, ACC40, Q15, ACC40, 1024, "SHIFT_REG", n_channels=2, layout="interleaved")
    f4.load(h4)
    assert f4.path == "fir_ovs"
    x = rng.integers(-32768, 32767, size=(160000, 2), endpoint=True).astype(np.int16)
    parts = [f4.run(x[:100]), f4.run(x[100:20000])]
    assert f4.ovs_margin()[1] == 0.0                # the residual monitor has seen no overlap-save launch so far
    parts += [f4.run(x[20000:130000])]
    assert 0 < f4.ovs_margin()[1] < 0.01            # ... and now it has
    parts += [f4.run(x[130000:130700]), f4.run(x[130700:])]
    y = np.concatenate(parts)
    for c in range(2):
        assert np.array_equal(y[:, c], ofir(oracle, Q15, Q15, ACC40, ACC40, 1024, "SHIFT_REG", h4, x[:, c])), c
    # fewer than 96 taps, formats outside the q15 family, order-dependent accumulators: never
    assert engine.ac_fir_const_coeffs(Q15, ACC40, Q15, ACC40, 95, "SHIFT_REG", h[:95]).path == "fir_q15"
    assert engine.ac_fir_const_coeffs(Q15, ACC40, Q15, ACC40, 96, "SHIFT_REG", h[:96]).path == "fir_ovs"
    assert engine.ac_fir_const_coeffs((20, 5), ACC40, Q15, ACC40, 128, "SHIFT_REG", h[:128]).path == "fir_q24"
    assert engine.ac_fir_const_coeffs(Q15, Q15, Q15, (24, 4, True, "AC_TRN", "AC_SAT"), 128, "SHIFT_REG", h[:128]).path == "fir_generic"
    # taps whose 1-norm pushes the error bound past 1/2 are refused (2048 taps at full scale)
    big = np.full(2048, -32768, dtype=np.int64)
    g = engine.ac_fir_const_coeffs(Q15, ACC40, Q15, ACC40, 2048, "SHIFT_REG", big)
    assert g.path == "fir_q15" and g.ovs_margin()[0] > 0.49
    monkeypatch.setenv("B2D_FIR_OVS", "0")
    assert engine.ac_fir_const_coeffs(Q15, ACC40, Q15, ACC40, 256, "SHIFT_REG", h).path == "fir_q15"


@pytest.mark.gpu
@pytest.mark.parametrize("ft", ["SHIFT_REG", "ROTATE_SHIFT", "C_BUFF", "FOLD_EVEN", "FOLD_ODD", "TRANSPOSED"])
@pytest.mark.parametrize("taps", [96, 255, 256, 257, 700, 1024])
def test_every_architecture_chunked_with_reload(engine, oracle, forced, ft, taps):
    if (ft == "FOLD_EVEN" and taps % 2) or (ft == "FOLD_ODD" and taps % 2 == 0):
        pytest.skip("fold parity")
    rng = np.random.default_rng(taps * 31 + len(ft))
    n = 9000
    x = oracle.rand_raw(rng, Q15, n)
    h1, h2 = oracle.rand_raw(rng, Q15, taps), oracle.rand_raw(rng, (14, 1), taps)
    b = oracle.FirB(Q15, Q15, ACC40, ACC40, taps, ft)
    f = engine.ac_fir_load_coeffs(Q15, ACC40, Q15, ACC40, taps, ft)
    want, got = [], []
    for k, (lo, hi) in enumerate(((0, 1), (1, 4000), (4000, 4003), (4003, 8000), (8000, n))):
        if k in (0, 3):
            hh = h1 if k == 0 else h2
            b.load(hh)
            f.load(hh)
        want.append(b.run(x[lo:hi]))
        got.append(np.atleast_1d(f.run(x[lo:hi].astype(np.int16))))
    assert f.path == "fir_ovs"
    assert np.array_equal(np.concatenate(got).astype(np.int64), np.concatenate(want)), (ft, taps)
    assert resid_ok(f, forced), f.ovs_margin()


@pytest.mark.gpu
@pytest.mark.parametrize("fmts", [
    ((16, 1), (16, 1), (40, 8), (18, 2, True, "AC_RND", "AC_SAT")),          # converting epilogue
    ((16, 0, False), (16, 0, False), (48, 16, False), (48, 16, False)),      # unsigned samples and taps
    ((16, 1), (12, 4, False), (36, 9), (36, 9)),                             # signed x unsigned, narrower accumulator (wraps)
    ((12, 0, False), (14, 2), (30, 6), (20, 4)),                             # left shift into the accumulator
    ((16, 1), (16, 1), (33, 1), (33, 1)),                                    # accumulator narrower than the sums: wraps
])
@pytest.mark.parametrize("layout,C", [("interleaved", 2), ("planar", 3), ("planar", 1)])
def test_formats_layouts_and_extremes(engine, oracle, forced, fmts, layout, C):
    fi, fc, fa, fo = fmts
    import zlib
    rng = np.random.default_rng(zlib.crc32(str((fmts, layout, C)).encode()))
    taps, n = 160, 9001
    for kind in ("uniform", "min", "alt"):
        x = np.stack([oracle.rand_raw(rng, fi, n, kind) for _ in range(C)])
        hs = [oracle.rand_raw(rng, fc, taps, "min" if kind == "min" else "uniform") for _ in range(C)]
        if layout == "interleaved":
            hs = [hs[0]] * C
        f = engine.ac_fir_prog_coeffs(fi, fo, fc, fa, taps, "SHIFT_REG", n_channels=C, layout=layout)
        for c in range(C):
            f.load(hs[c], channel=c)
        assert f.path == "fir_ovs", (fmts, f.ovs_margin())
        xin = x.astype(np.int16 if fi[0] <= 16 else np.int32)
        xin = xin[0] if C == 1 else (np.ascontiguousarray(xin.T) if layout == "interleaved" else xin)
        planar = layout == "planar" and C > 1
        parts = [f.run(xin[:, :5000] if planar else xin[:5000]), f.run(xin[:, 5000:] if planar else xin[5000:])]
        y = np.concatenate([np.atleast_1d(p_) for p_ in parts], axis=1 if planar else 0)
        y = y.reshape(1, -1) if C == 1 else (y.T if layout == "interleaved" else y)
        for c in range(C):
            assert np.array_equal(y[c].astype(np.int64), ofir(oracle, fi, fc, fa, fo, taps, "SHIFT_REG", hs[c], x[c])), (fmts, layout, c, kind)
        assert resid_ok(f, forced), f.ovs_margin()


@pytest.mark.gpu
def test_device_buffers_unaligned_views_and_state(engine, oracle, forced):
    import torch
    rng = np.random.default_rng(5)
    taps, n = 300, 30011
    x = rng.integers(-32768, 32767, size=(n, 2), endpoint=True).astype(np.int16)
    h = oracle.rand_raw(rng, Q15, taps)
    want = np.stack([ofir(oracle, Q15, Q15, ACC40, ACC40, taps, "SHIFT_REG", h, x[:, c]) for c in range(2)], axis=1)
    f = engine.ac_fir_load_coeffs(Q15, ACC40, Q15, ACC40, taps, "SHIFT_REG", n_channels=2, layout="interleaved")
    f.load(h)
    xd = torch.from_numpy(x).cuda()
    y1 = f.run(xd[:7])                              # the next view starts 28 bytes into the allocation
    y2 = f.run(xd[7:20001])
    blob = f.get_state()
    y3 = f.run(xd[20001:])
    torch.cuda.synchronize()
    assert np.array_equal(torch.cat([y1, y2, y3]).cpu().numpy(), want)
    g = engine.ac_fir_load_coeffs(Q15, ACC40, Q15, ACC40, taps, "SHIFT_REG", n_channels=2, layout="interleaved")
    g.load(h)
    g.set_state(blob)
    out = torch.empty((n - 20001 + 1, 2), dtype=torch.int64, device="cuda")
    assert np.array_equal(g.run(xd[20001:], out=out.reshape(-1)[2:]).cpu().numpy(), want[20001:])   # output 16 bytes in: still aligned
    assert g.path == "fir_ovs"


FUZZ_SEED = int(os.environ.get("B2D_FUZZ_SEED", "0"))
Q_MODES = ["AC_TRN", "AC_RND", "AC_TRN_ZERO", "AC_RND_ZERO", "AC_RND_INF", "AC_RND_MIN_INF", "AC_RND_CONV", "AC_RND_CONV_ODD"]
O_MODES = ["AC_WRAP", "AC_SAT", "AC_SAT_ZERO", "AC_SAT_SYM"]


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(12))
def test_random_q15_family_draws(engine, oracle, forced, i):
    """Random members of the format family the path serves (samples and taps of up to 16 bits, signed or not, an
    accumulator that takes the products without dropping bits and wraps, any OUT_TYPE incl. every Q / O mode), random tap
    counts 96..2048, layouts, chunkings and a coefficient change: engine == Oracle B.  B2D_FUZZ_SEED draws again."""
    rng = np.random.default_rng(FUZZ_SEED * 1000 + 77 + i)
    Wi, Wc = int(rng.integers(2, 17)), int(rng.integers(2, 17))
    fi = (Wi, int(rng.integers(-2, Wi + 3)), bool(rng.integers(0, 2)))
    fc = (Wc, int(rng.integers(-2, Wc + 3)), bool(rng.integers(0, 2)))
    Fp = (fi[0] - fi[1]) + (fc[0] - fc[1])
    Fa = Fp + int(rng.integers(0, 5))                                        # exact left shift into the accumulator
    Wa = int(rng.integers(12, 65))
    fa = (Wa, Wa - Fa, bool(rng.integers(0, 4)), ["AC_TRN", "AC_RND"][int(rng.integers(0, 2))], "AC_WRAP")
    Wo = int(rng.integers(4, 65))
    fo = (Wo, Wo - (Fa - int(rng.integers(0, 10))), bool(rng.integers(0, 2)), Q_MODES[int(rng.integers(0, 8))], O_MODES[int(rng.integers(0, 4))])
    if i % 3 == 0:
        fo = fa
    taps = int(rng.choice([96, 100, 128, 255, 256, 257, 511, 512, 513, 1000, 1024, 1500, 2048]))
    ft = ["SHIFT_REG", "C_BUFF", "TRANSPOSED", "FOLD_EVEN", "FOLD_ODD"][int(rng.integers(0, 5))]
    if ft == "FOLD_EVEN" and taps % 2:
        taps += 1
    if ft == "FOLD_ODD" and taps % 2 == 0:
        taps -= 1
    if ft == "FOLD_ODD" and not (fa[2] and Fa >= fi[0] - fi[1] and fi[0] + 2 + (Fa - (fi[0] - fi[1])) <= fa[0]):
        ft = "SHIFT_REG"                                                     # the pre-add must be exact in ACC_TYPE for the q15 family
    layout, C = [("interleaved", 2), ("planar", 1), ("planar", 2), ("planar", 5)][int(rng.integers(0, 4))]
    n = int(rng.integers(1, 4)) * 4096 + int(rng.integers(0, 4096))
    x = np.stack([oracle.rand_raw(rng, fi, n) for _ in range(C)])
    hs1 = [oracle.rand_raw(rng, fc, taps) for _ in range(C)]
    hs2 = [oracle.rand_raw(rng, fc, taps) for _ in range(C)]
    if layout == "interleaved":
        hs1, hs2 = [hs1[0]] * C, [hs2[0]] * C
    f = engine.ac_fir_load_coeffs(fi, fo, fc, fa, taps, ft, n_channels=C, layout=layout)
    obs = [oracle.FirB(fi, fc, fa, fo, taps, ft) for _ in range(C)]
    cut1, cut2 = int(rng.integers(1, n // 2)), int(rng.integers(n // 2, n))
    want = [[] for _ in range(C)]
    got = []
    for k, (lo, hi) in enumerate(((0, cut1), (cut1, cut2), (cut2, n))):
        if k in (0, 2):
            for c in range(C):
                hh = (hs1 if k == 0 else hs2)[c]
                obs[c].load(hh)
                f.load(hh, channel=c)
        for c in range(C):
            want[c].append(obs[c].run(x[c, lo:hi]))
        seg = x[:, lo:hi].astype(np.int16)
        seg = seg[0] if C == 1 else (np.ascontiguousarray(seg.T) if layout == "interleaved" else np.ascontiguousarray(seg))
        y = f.run(seg)
        got.append(y.reshape(1, -1) if C == 1 else (y.T if layout == "interleaved" else y))
    bound, resid = f.ovs_margin()
    info = (fi, fc, fa, fo, taps, ft, layout, C, n, f.path, bound, resid)
    if f.path == "fir_ovs":
        assert resid_ok(f, forced), info
    else:
        assert f.path == "fir_q15" and bound >= 0.49, info                    # refused by the error bound only
    y = np.concatenate(got, axis=1).astype(np.int64)
    for c in range(C):
        assert np.array_equal(y[c], np.concatenate(want[c])), (info, c)
