"""Engine == Oracle B on RANDOM formats (GPU; part of the default `-m gpu` run).

tests/test_oracle_fuzz.py pins the restatement to the unmodified reference templates on drawn instantiations; this is
the same draw through the CUDA engine: formats outside the hand-picked tables, all 8 x 4 accumulator / output modes
(the generic kernels), whatever fast path the predicates select, chunked calls, coefficient changes mid-stream (every
architecture, TRANSPOSED included).  First run on a B200 at the start of round 2 with six seeds (B2D_FUZZ_SEED=0..5,
profiles/r02_upfir_ab_and_engine_fuzz.txt): all green; B2D_FUZZ_SEED draws again.
"""
import numpy as np
import pytest

import test_oracle_fuzz as F

pytestmark = [pytest.mark.gpu]
SEED = F.SEED


def make_fir(E, kind, fi, fc, fa, fo, taps, ft, coeffs):
    if kind == 0:
        return E.ac_fir_const_coeffs(fi, fo, fc, fa, taps, ft, coeffs)
    if kind == 1:
        f = E.ac_fir_load_coeffs(fi, fo, fc, fa, taps, ft)
        f.run(None, coeffs, True)
        return f
    f = E.ac_fir_prog_coeffs(fi, fo, fc, fa, taps, ft)
    f.load(coeffs)
    return f


@pytest.mark.parametrize("i", range(16))
def test_fir_random_formats(engine, oracle, i):
    rng = np.random.default_rng(SEED + 9000 + i)
    fi, fc, fa, fo, nt = F.draw_fir(rng, 1)[0]
    x = oracle.rand_raw(rng, fi, 3000)
    h, h2 = oracle.rand_raw(rng, fc, nt), oracle.rand_raw(rng, fc, nt)
    cuts = ((0, 1), (1, nt + 3), (nt + 3, nt + 5), (nt + 5, 3000))
    for ft in ("SHIFT_REG", "ROTATE_SHIFT", "C_BUFF", "FOLD_EVEN", "FOLD_ODD", "TRANSPOSED"):
        if (ft == "FOLD_EVEN" and nt % 2) or (ft == "FOLD_ODD" and nt % 2 == 0):
            continue
        reload_ = i % 3 != 0                       # load / prog classes: other taps mid-stream (twice, 2 samples apart)
        b = oracle.FirB(fi, fc, fa, fo, nt, ft)
        f = make_fir(engine, i % 3, fi, fc, fa, fo, nt, ft, h)
        b.load(h)
        want, got = [], []
        for k, (a, c) in enumerate(cuts):
            if reload_ and k in (2, 3):
                hh = h2 if k == 2 else h
                b.load(hh)
                f.load(hh)
            want.append(b.run(x[a:c]))
            got.append(np.atleast_1d(f.run(x[a:c])))
        assert np.array_equal(np.concatenate(got).astype(np.int64), np.concatenate(want)), ((fi, fc, fa, fo, nt), ft, f.path, reload_)


@pytest.mark.parametrize("mode", ["dec", "intr"])
@pytest.mark.parametrize("i", range(10))
def test_cic_random_formats(engine, oracle, mode, i):
    rng = np.random.default_rng(SEED + 9100 + i)
    R, M, N, fi, fo = F.draw_cic(rng, 1, mode)[0]
    if N > 16 or N * M > 64:
        pytest.skip("outside the engine's N / N*M limits")
    x = oracle.rand_raw(rng, fi, 4000 if mode == "dec" else 700)
    want = oracle.CicB(mode, fi, fo, R, M, N).run(x)
    f = (engine.ac_cic_dec_full if mode == "dec" else engine.ac_cic_intr_full)(fi, fo, R, M, N)
    y = np.concatenate([f.run(x[a:c]) for a, c in ((0, 1), (1, 58), (58, 61), (61, x.size))])
    assert np.array_equal(y.astype(np.int64), want), ((R, M, N, fi, fo), f.path)


@pytest.mark.parametrize("i", range(10))
def test_poly_dec_random_formats(engine, oracle, i):
    rng = np.random.default_rng(SEED + 9200 + i)
    fi, fc, fa, fo = F.draw_mac_formats(rng)
    nt, df = int(rng.integers(1, 13)), int(rng.integers(2, 7))
    x, h = oracle.rand_raw(rng, fi, 300 * df + 3), oracle.rand_raw(rng, fc, nt * df)
    b = oracle.PdB(fi, fc, fa, fo, nt, df)
    b.load(h)
    want = b.run(x)
    f = engine.ac_poly_dec(fi, fc, fa, fo, nt, df, coeffs=h)
    y = np.concatenate([f.run(x[a:c]) for a, c in ((0, 1), (1, df + 2), (df + 2, x.size))])
    assert np.array_equal(y.astype(np.int64), want), ((fi, fc, fa, fo, nt, df), f.path)


@pytest.mark.parametrize("i", range(12))
def test_poly_intr_random_formats(engine, oracle, i):
    rng = np.random.default_rng(SEED + 9300 + i)
    fi, fc, fa, fo = F.draw_mac_formats(rng)
    ft = ["FOLD_EVEN", "FOLD_ODD", "FOLD_ANTI"][i % 3]
    nt = int(rng.integers(1, 7)) * 2 if ft == "FOLD_EVEN" else (int(rng.integers(0, 6)) * 2 + 1 if ft == "FOLD_ODD" else int(rng.integers(1, 12)))
    IF = int(rng.integers(1, 6))
    from oracle import ref_configs as rc
    csz = rc.pi_coeffsz((fi, fc, fa, fo, nt, IF, ft))
    x, h = oracle.rand_raw(rng, fi, 900), oracle.rand_raw(rng, fc, csz)
    sign, corr = rng.integers(0, 2, IF), rng.integers(0, IF, IF)
    b = oracle.PiB(fi, fc, fa, fo, nt, IF, ft)
    b.load(h, sign, corr)
    want = b.run(x)
    f = engine.ac_poly_intr(fi, fc, fa, fo, nt, IF, ft, coeffs=h, sign=sign, corr=corr)
    y = np.concatenate([np.asarray(f.run(x[a:c])).reshape(-1) for a, c in ((0, 1), (1, 9), (9, 900))])
    assert np.array_equal(y.astype(np.int64), want), ((fi, fc, fa, fo, nt, IF, ft), f.path)


@pytest.mark.parametrize("i", range(8))
def test_intg_dump_random_formats(engine, oracle, i):
    rng = np.random.default_rng(SEED + 9400 + i)
    fi = F.rand_fmt(rng, 2, 32)
    Fa = (fi[0] - fi[1]) + int(rng.integers(-8, 5))
    Wa, Wo = int(rng.integers(8, 65)), int(rng.integers(4, 65))
    fa = (Wa, Wa - Fa, bool(rng.integers(0, 4)), F.Q_MODES[int(rng.integers(0, 8))], F.O_MODES[int(rng.integers(0, 4))])
    fo = (Wo, Wo - (Fa - int(rng.integers(0, 8))), bool(rng.integers(0, 2)), F.Q_MODES[int(rng.integers(0, 8))], F.O_MODES[int(rng.integers(0, 4))])
    NS, CHN = int(rng.integers(4, 65)), int(rng.integers(1, 6))
    tok = rng.integers(1, NS + 1, 40)
    tok[[4, 11, 20]] = [0, NS + 5, NS]
    n = int(sum(oracle.id_frame_samples(v, NS, CHN) for v in tok))
    x = oracle.rand_raw(rng, fi, n)
    want = np.asarray(oracle.IdB(fi, fa, fo, NS, CHN).run(x, tok)).reshape(-1)
    y = np.asarray(engine.ac_intg_dump(fi, fa, fo, NS, CHN).run(x, tok)).reshape(-1)
    assert np.array_equal(y.astype(np.int64), want), ((fi, fa, fo, NS, CHN),)


@pytest.mark.parametrize("i", range(9))
def test_mv_avg_random_formats(engine, oracle, i):
    rng = np.random.default_rng(SEED + 9500 + i)
    maxs, taps, wt, fi, fo, fa, fc = F.draw_mv(rng, 3)[i % 3]
    c = oracle.rand_raw(rng, fc, taps)
    f = engine.ac_mv_avg(maxs, taps, wt, fi, fo, fa, fc, c)
    for ns in (taps, maxs):
        x = oracle.rand_raw(rng, fi, 5 * ns)
        assert np.array_equal(f.run(x, ns).astype(np.int64), oracle.mv_run_b(fi, fo, fa, fc, taps, wt, c, x, ns)), ((maxs, taps, wt, fi, fo, fa, fc), ns)
