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
mid-stream); the engine, which carries input history only, refuses such a change instead of approximating it.
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


@pytest.mark.gpu
def test_engine_refuses_a_coefficient_change_on_a_running_transposed_filter(engine, oracle):
    rng = np.random.default_rng(6)
    taps = 16
    x = oracle.rand_raw(rng, Q15, 300).astype(np.int16)
    h1, h2 = oracle.rand_raw(rng, Q15, taps), oracle.rand_raw(rng, Q15, taps)
    ob = oracle.FirB(Q15, Q15, ACC40, ACC40, taps, "TRANSPOSED")
    ob.load(h1)
    want = ob.run(x)
    f = engine.ac_fir_prog_coeffs(Q15, ACC40, Q15, ACC40, taps, "TRANSPOSED")
    y = np.concatenate([f.run(x[:100], h1), f.run(x[100:200], h1)])          # the same array on every call: fine
    assert np.array_equal(y.astype(np.int64), want[:200])
    with pytest.raises(engine.B2dError) as e:
        f.run(x[200:], h2)
    assert e.value.status == -1                                              # B2D_EUNSUPPORTED
    assert np.array_equal(f.run(x[200:], h1).astype(np.int64), want[200:])   # the refusal left the filter untouched
    f.reset()
    ob2 = oracle.FirB(Q15, Q15, ACC40, ACC40, taps, "TRANSPOSED")
    ob2.load(h2)
    assert np.array_equal(f.run(x, h2).astype(np.int64), ob2.run(x))         # after reset() any taps are welcome
    g = engine.ac_fir_prog_coeffs(Q15, ACC40, Q15, ACC40, taps, "SHIFT_REG")  # other architectures: the delay line is the state
    ob3 = oracle.FirB(Q15, Q15, ACC40, ACC40, taps, "SHIFT_REG")
    ob3.load(h1)
    w = ob3.run(x[:100])
    ob3.load(h2)
    w = np.concatenate([w, ob3.run(x[100:])])
    assert np.array_equal(np.concatenate([g.run(x[:100], h1), g.run(x[100:], h2)]).astype(np.int64), w)
