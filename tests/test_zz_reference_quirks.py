"""Two places where "what the reference computes" is not what its structure suggests (runs last: the file name sorts
after every other test).

1. The comb for differential delays M > 2.

ac_cic_full_core.h:247-251 shifts comb_dly_ln with an ascending copy loop, so the delay the reference really applies is
min(M, 2) while its lossless width still grows with M.  tests/test_oracle_fuzz.py found the difference between the
unmodified templates and the first restatement; tests/golden/cic_comb_quirk.npz (make_comb_quirk_golden.py) holds the
real reference's outputs for 14 such instantiations.  CPU: the restatement reproduces them.  GPU: the engine does
(generic kernel at the M-wide internal type with delay 2; the cascade falls back to its two-stage path).

2. TRANSPOSED keeps ACC_TYPE partial sums (reg_trans[], ac_fir_load_coeffs.h:265-278): after a coefficient change its next
N_TAPS-1 outputs mix old-tap partial sums with new-tap products.  The restatement models that (the random sweep reloads
mid-stream, pinning it to the real class); the engine, which carries input history, converts the history into those
pending partial sums at the change (round 1 refused the change instead).
"""
import numpy as np
import pytest

from conftest import golden

Q15, ACC40 = (16, 1), (40, 8)


def cases():
    g = golden("cic_comb_quirk.npz")
    out = []
    for k in range(int(g["n"][0])):
        fi, fo = (tuple(int(v) for v in g[f"c{k}_{n}"]) for n in ("fin", "fout"))
        R, M, N = (int(v) for v in g[f"c{k}_rmn"])
        out.append(("intr" if int(g[f"c{k}_mode"][0]) else "dec", R, M, N, fi, fo, g[f"c{k}_x"], g[f"c{k}_y"], list(g[f"c{k}_counts"])))
    return out


CASES = cases()
IDS = [f"{c[0]}-R{c[1]}M{c[2]}N{c[3]}-out{c[5][0]}" for c in CASES]
CUTS = ((0, 1), (1, 10), (10, 13), (13, None))


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_restatement_reproduces_the_reference_for_m_above_2(oracle, case):
    mode, R, M, N, fi, fo, x, y, counts = case
    assert M > 2
    b = oracle.CicB(mode, fi, fo, R, M, N)
    parts = [b.run(x[lo:hi]) for lo, hi in CUTS]
    assert [p.size for p in parts] == counts
    assert np.array_equal(np.concatenate(parts), y)
    # ... and it is the M = 2 filter at the wider type, not the M-delay one
    assert np.array_equal(oracle.CicB(mode, fi, fo, R, 2, N).run(x), y)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_engine_reproduces_the_reference_for_m_above_2(engine, case):
    mode, R, M, N, fi, fo, x, y, counts = case
    cls = engine.ac_cic_dec_full if mode == "dec" else engine.ac_cic_intr_full
    f = cls(fi, fo, R, M, N)
    assert f.path == "cic_generic"
    parts = [f.run(x[lo:hi]) for lo, hi in CUTS]
    assert [p.size for p in parts] == counts, (mode, R, M, N)
    assert np.array_equal(np.concatenate(parts).astype(np.int64), y), (mode, R, M, N)
    f.reset()
    assert np.array_equal(f.run(x).astype(np.int64), y)


@pytest.mark.gpu
def test_cascade_with_m_above_2_takes_the_two_stage_path(engine, oracle):
    rng = np.random.default_rng(3)
    R, M, N, taps = 4, 3, 3, 31
    W = oracle.cic_int_width("intr", Q15, R, M, N)
    mid = (W, W - 15)
    x = oracle.rand_raw(rng, Q15, 5003)
    h = oracle.rand_raw(rng, Q15, taps)
    f = engine.cic_intr_fir_cascade(Q15, mid, R, M, N, ACC40, Q15, ACC40, taps, "SHIFT_REG", coeffs=h)
    assert f.path == "cicfir_two_stage"
    fir = oracle.FirB(mid, Q15, ACC40, ACC40, taps, "SHIFT_REG")
    fir.load(h)
    want = fir.run(oracle.CicB("intr", Q15, mid, R, M, N).run(x))
    y = np.concatenate([f.run(x[:7]), f.run(x[7:8]), f.run(x[8:])])
    assert np.array_equal(y.astype(np.int64), want)


# ------------------------------------------------------------------------------ TRANSPOSED and coefficient changes
def test_transposed_reload_differs_from_the_direct_form_in_the_reference(oracle):
    """The behaviour the engine refuses to approximate, shown on the restatement (pinned to the real class by
    tests/test_oracle_fuzz.py): N_TAPS-1 outputs after the change differ from the direct form, then both agree again."""
    rng = np.random.default_rng(5)
    taps = 16
    x = oracle.rand_raw(rng, Q15, 200)
    h1, h2 = oracle.rand_raw(rng, Q15, taps), oracle.rand_raw(rng, Q15, taps)
    outs = {}
    for ft in ("SHIFT_REG", "TRANSPOSED"):
        f = oracle.FirB(Q15, Q15, ACC40, ACC40, taps, ft)
        f.load(h1)
        a = f.run(x[:100])
        f.load(h2)
        outs[ft] = np.concatenate([a, f.run(x[100:])])
    d = np.flatnonzero(outs["SHIFT_REG"] != outs["TRANSPOSED"])
    assert d.size and d.min() >= 100 and d.max() <= 100 + taps - 2
    # y[n] = sum_i h_{set active when x[n-i] arrived}[i] * x[n-i]
    n = 105
    want = sum(int((h2 if n - i >= 100 else h1)[i]) * int(x[n - i]) for i in range(taps)) << 2
    want = (want + (1 << 39)) % (1 << 40) - (1 << 39)
    assert int(outs["TRANSPOSED"][n]) == want


TR_FORMATS = [
    # (in, coeff, acc, out, taps, expected path): the three kernel families and order-dependent accumulators
    (Q15, Q15, ACC40, ACC40, 16, "fir_q15"),
    (Q15, Q15, ACC40, ACC40, 256, "fir_q15"),
    ((16, 1), (16, 1), (24, 4), (16, 1), 31, "fir_wide"),                                     # per-tap truncation, narrow output
    ((28, 6), (23, 7), (64, 32), (64, 32), 27, "fir_wide"),                                   # the reference prog bench's formats
    ((16, 1), (16, 1), (24, 4, True, "AC_TRN", "AC_SAT"), (16, 1, True, "AC_RND_CONV", "AC_SAT_SYM"), 9, "fir_generic"),
    ((16, 1), (16, 1), (30, 6, True, "AC_TRN_ZERO", "AC_SAT_ZERO"), (12, 1, True, "AC_RND_INF", "AC_SAT"), 10, "fir_generic"),
    ((12, 0, False), (14, 2), (30, 6, False, "AC_RND", "AC_WRAP"), (20, 4), 12, None),
]


def _transposed_want(oracle, fmt, x, sets, cuts):
    fi, fc, fa, fo, taps, _ = fmt
    ob = oracle.FirB(fi, fc, fa, fo, taps, "TRANSPOSED")
    parts = []
    for (lo, hi), h in zip(cuts, sets):
        ob.load(h)
        parts.append(ob.run(x[lo:hi]))
    return np.concatenate(parts)


@pytest.mark.gpu
@pytest.mark.parametrize("cls", ["load", "prog"])
@pytest.mark.parametrize("fmt", TR_FORMATS, ids=[f"{f[5]}-{f[4]}" for f in TR_FORMATS])
def test_engine_follows_transposed_partial_sums_across_coefficient_changes(engine, oracle, fmt, cls):
    """ac_fir_load_coeffs.h:265-278 / ac_fir_prog_coeffs.h:232-247: y[n] = sum_i h_{set active when x[n-i] arrived}[i] x[n-i],
    accumulated oldest sample first.  Changes closer together than N_TAPS-1 samples, a change after a single sample,
    a re-load of equal taps and chunked calls in between."""
    fi, fc, fa, fo, taps, path = fmt
    rng = np.random.default_rng(60 + taps)
    x = oracle.rand_raw(rng, fi, 3 * taps + 700)
    hs = [oracle.rand_raw(rng, fc, taps) for _ in range(4)]
    sets = [hs[0], hs[1], hs[2], hs[2], hs[3], hs[0]]
    edges = [0, 2 * taps + 5, 2 * taps + 6, 2 * taps + 6 + max(taps // 2, 1), 3 * taps + 40, 3 * taps + 300, x.size]
    cuts = list(zip(edges[:-1], edges[1:]))
    want = _transposed_want(oracle, fmt, x, sets, cuts)
    if cls == "load":
        f = engine.ac_fir_load_coeffs(fi, fo, fc, fa, taps, "TRANSPOSED")
        run = lambda seg, h: f.run(seg, h, True)
    else:
        f = engine.ac_fir_prog_coeffs(fi, fo, fc, fa, taps, "TRANSPOSED")
        run = lambda seg, h: f.run(seg, h)
    if path:
        assert f.path == path
    parts = []
    for (lo, hi), h in zip(cuts, sets):
        mid = lo + (hi - lo) // 3
        parts.append(np.atleast_1d(run(x[lo:mid], h)) if mid > lo else np.empty(0, dtype=np.int64))
        parts.append(np.atleast_1d(run(x[mid:hi], h)))
    y = np.concatenate([np.asarray(p).astype(np.int64) for p in parts])
    bad = np.flatnonzero(y != want)
    assert bad.size == 0, (fmt, cls, bad[:8])
    # the direct form takes the new taps at once: same stream, different outputs after each change -- and still the oracle's
    g = engine.ac_fir_prog_coeffs(fi, fo, fc, fa, taps, "SHIFT_REG")
    od = oracle.FirB(fi, fc, fa, fo, taps, "SHIFT_REG")
    yd, wd = [], []
    for (lo, hi), h in zip(cuts, sets):
        yd.append(np.atleast_1d(g.run(x[lo:hi], h)))
        od.load(h)
        wd.append(od.run(x[lo:hi]))
    assert np.array_equal(np.concatenate(yd).astype(np.int64), np.concatenate(wd))


@pytest.mark.gpu
def test_transposed_change_per_sample_and_checkpoint(engine, oracle):
    """ac_fir_prog_coeffs with adaptive taps: another array on EVERY call (one sample per call, as the reference's run()),
    a checkpoint taken while partial sums are pending, and per-channel changes on a two-channel handle."""
    rng = np.random.default_rng(77)
    taps = 12
    x = oracle.rand_raw(rng, Q15, 90).astype(np.int16)
    hs = [oracle.rand_raw(rng, Q15, taps) for _ in range(x.size)]
    ob = oracle.FirB(Q15, Q15, ACC40, ACC40, taps, "TRANSPOSED")
    want = []
    for k in range(x.size):
        ob.load(hs[k])
        want.append(ob.run(x[k:k + 1])[0])
    want = np.array(want)
    f = engine.ac_fir_prog_coeffs(Q15, ACC40, Q15, ACC40, taps, "TRANSPOSED")
    y = [int(np.atleast_1d(f.run(x[k:k + 1], hs[k]))[0]) for k in range(40)]
    blob = f.get_state()
    y += [int(np.atleast_1d(f.run(x[k:k + 1], hs[k]))[0]) for k in range(40, x.size)]
    assert np.array_equal(np.array(y), want)
    g = engine.ac_fir_prog_coeffs(Q15, ACC40, Q15, ACC40, taps, "TRANSPOSED")
    g.load(hs[39])
    g.set_state(blob)
    y2 = [int(np.atleast_1d(g.run(x[k:k + 1], hs[k]))[0]) for k in range(40, x.size)]
    assert np.array_equal(np.array(y2), want[40:])
    f.reset()                                                                 # reset() drops the pending sums as well
    ob2 = oracle.FirB(Q15, Q15, ACC40, ACC40, taps, "TRANSPOSED")
    ob2.load(hs[3])
    assert np.array_equal(np.atleast_1d(f.run(x, hs[3])).astype(np.int64), ob2.run(x))
    # two channels, the taps of channel 1 change while channel 0 keeps its set
    xx = oracle.rand_raw(rng, Q15, 2 * 300).astype(np.int16).reshape(2, 300)
    m = engine.ac_fir_load_coeffs(Q15, ACC40, Q15, ACC40, taps, "TRANSPOSED", n_channels=2)
    m.load(hs[0])
    parts = [m.run(xx[:, :100])]
    m.load(hs[1], channel=1)
    parts.append(m.run(xx[:, 100:105]))
    m.load(hs[2], channel=1)
    parts.append(m.run(xx[:, 105:]))
    ym = np.concatenate(parts, axis=1).astype(np.int64)
    o0 = oracle.FirB(Q15, Q15, ACC40, ACC40, taps, "TRANSPOSED")
    o0.load(hs[0])
    assert np.array_equal(ym[0], o0.run(xx[0]))
    o1 = oracle.FirB(Q15, Q15, ACC40, ACC40, taps, "TRANSPOSED")
    w1 = []
    for (lo, hi), h in zip(((0, 100), (100, 105), (105, 300)), (hs[0], hs[1], hs[2])):
        o1.load(h)
        w1.append(o1.run(xx[1, lo:hi]))
    assert np.array_equal(ym[1], np.concatenate(w1))
