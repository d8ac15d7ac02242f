"""CPU suite: pins the oracle (test infrastructure) against every vector the reference holds for the hot path.

* Oracle B (oracle/oracle_b.c, header-free integer restatement) vs the reference's two bit-exact CIC golden
  vectors (tests/ac_cic_{dec,intr}_full_{input,ref}.txt, committed as raw integers in tests/golden/).
* Oracle B vs the outputs of the UNMODIFIED reference classes (Oracle A, compiled from /root/reference over the
  clean-room ac_types shim) on seeded random inputs for every configuration of oracle/ref_configs.py x ftype,
  committed in tests/golden/ref_outputs.npz -- and live A == B sweeps when oracle/_ref/libacdsp_ref.so exists.
* The three FIR benches: stimulus, coefficients, the reference class's own output and its SQNR vs the MATLAB doubles.
* Known-answer tests derived from the semantics (marked DERIVED: not from the reference's tests).
"""
import math

import numpy as np
import pytest

from conftest import golden
from oracle import ref_configs as rc


def test_cic_dec_golden(oracle):
    g = golden("cic_dec_golden.npz")
    f = oracle.CicB("dec", (32, 16), (48, 32), int(g["R"]), int(g["M"]), int(g["N"]))
    y = f.run(g["x"])
    assert y.size == 1430                                   # rtest_ac_cic_dec_full.cpp: 10004 in -> 1430 out
    assert np.array_equal(y[:1429], g["ref"])               # first 1429 compared with diff == 0 (:129-134)


def test_cic_intr_golden(oracle):
    g = golden("cic_intr_golden.npz")
    f = oracle.CicB("intr", (32, 16), (49, 33), int(g["R"]), int(g["M"]), int(g["N"]))
    y = f.run(g["x"])
    assert y.size == 6990 and np.array_equal(y, g["ref"])   # rtest_ac_cic_intr_full.cpp:99,123-133


@pytest.mark.parametrize("cls", ["const", "load", "prog"])
def test_fir_bench(oracle, cls):
    g = golden(f"fir_bench_{cls}.npz")
    fi, fc, fa, fo = (tuple(int(v) for v in g[k]) for k in ("fin", "fcoeff", "facc", "fout"))
    f = oracle.FirB(fi, fc, fa, fo, int(g["taps"]), "FOLD_ODD")
    f.load(g["coeffs"])
    y = f.run(g["x"])
    assert np.array_equal(y, g["y"])                        # == the unmodified reference class's output
    F = fo[0] - fo[1]
    ref = g["ref_double"][: y.size]
    sqnr = 10 * math.log10(np.sum(ref * ref) / np.sum((y / float(1 << F) - ref) ** 2))
    assert abs(sqnr - float(g["sqnr"])) < 1e-9 and sqnr >= 60.0   # rtest_ac_fir_*_coeffs.cpp: SQNR >= 60 dB
    assert abs(sqnr - {"const": 84.2385, "load": 89.5576, "prog": 89.5576}[cls]) < 5e-4


@pytest.mark.parametrize("cfg", rc.fir_configs(), ids=lambda c: f"{c[1]}-{c[6]}")
def test_fir_vs_reference_outputs(oracle, ref_outputs, cfg):
    cid, _name, fi, fc, fa, fo, taps = cfg
    x = ref_outputs[f"fir{cid}_x"]
    for ft in oracle.FTYPES[:6]:
        c = ref_outputs[f"fir{cid}_csym" if ft.startswith("FOLD") else f"fir{cid}_c"]
        f = oracle.FirB(fi, fc, fa, fo, taps, ft)
        f.load(c)
        y = np.concatenate([f.run(x[:5]), f.run(x[5:6]), f.run(x[6:])])
        assert np.array_equal(y, ref_outputs[f"fir{cid}_{ft}_y"]), (cfg, ft)


def test_cic_vs_reference_outputs(oracle, ref_outputs):
    for cid, (mode, R, M, N, fi, fo) in enumerate(rc.CIC_CONFIGS):
        x = ref_outputs[f"cic{cid}_x"]
        f = oracle.CicB(mode, fi, fo, R, M, N)
        parts = [f.run(x[:1]), f.run(x[1:10]), f.run(x[10:13]), f.run(x[13:])]
        assert [p.size for p in parts] == list(ref_outputs[f"cic{cid}_counts"]), (mode, R, M, N)
        assert np.array_equal(np.concatenate(parts), ref_outputs[f"cic{cid}_y"]), (mode, R, M, N)


def test_live_reference_equals_restatement(oracle):
    """A == B on fresh random data (only where the reference could be compiled: the dev container)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libacdsp_ref.so not built (no /root/reference here)")
    rng = np.random.default_rng(7)
    for cid, _name, fi, fc, fa, fo, taps in rc.fir_configs()[::3]:
        x = oracle.rand_raw(rng, fi, 200)
        c = oracle.rand_raw(rng, fc, taps)
        for k, ft in enumerate(oracle.FTYPES[:6]):
            a = oracle.FirA(oracle.FIR_CLASSES[k % 3], fi, fc, fa, fo, taps, ft)
            b = oracle.FirB(fi, fc, fa, fo, taps, ft)
            a.load(c), b.load(c)
            assert np.array_equal(np.concatenate([a.run(x[:33]), a.run(x[33:])]), b.run(x))
    for mode, R, M, N, fi, fo in rc.CIC_CONFIGS[::5]:
        x = oracle.rand_raw(rng, fi, 300 if mode == "dec" else 60)
        assert np.array_equal(oracle.CicA(mode, fi, fo, R, M, N).run(x), oracle.CicB(mode, fi, fo, R, M, N).run(x))


# ------------------------------------------------------------------ DERIVED known-answer tests
def test_kat_fir_impulse_and_wrap(oracle):
    q15, acc = (16, 1), (40, 8)
    h = np.arange(1, 17, dtype=np.int64) * 1000
    f = oracle.FirB(q15, q15, acc, acc, 16, "SHIFT_REG")
    f.load(h)
    x = np.zeros(40, dtype=np.int64)
    x[3] = 1
    y = f.run(x)
    assert np.array_equal(y[3:19], h << 2)                 # s = 15 + 15 - 32 = -2: exact left shift by 2
    # all -1.0 x all -1.0 over 256 taps: sum = 256 * 2^30 = 2^38, << 2 = 2^40 -> wraps to 0 in <40,8>
    f = oracle.FirB(q15, q15, acc, acc, 256, "SHIFT_REG")
    f.load(np.full(256, -32768))
    y = f.run(np.full(300, -32768))
    assert y[255] == 0 and y[254] == (255 << 32) - (1 << 40)


def test_kat_cic_dc_gain_and_counts(oracle):
    R, M, N = 8, 1, 4
    y = oracle.CicB("dec", (16, 1), (28, 13), R, M, N).run(np.full(400, 5))
    assert y[-1] == 5 * (R * M) ** N                       # DC gain (RM)^N
    assert y.size == 50
    for K in (1, 2, 3, 20):
        n = oracle.CicB("intr", (16, 1), (20, 5), 4, 1, 3).run(np.arange(K)).size
        assert n == max(0, (K - 1) * 4 + 1 - 2)            # (K-1)R + 1 - (N-1)


def test_int_width(oracle):
    assert oracle.cic_int_width("dec", (16, 1), 8, 1, 4) == 28
    assert oracle.cic_int_width("dec", (16, 1), 8, 2, 4) == 32
    assert oracle.cic_int_width("intr", (16, 1), 4, 1, 3) == 20
    assert oracle.cic_int_width("dec", (32, 16), 7, 2, 4) == 48
    assert oracle.cic_int_width("intr", (32, 16), 7, 2, 5) == 49


# ------------------------------------------------------------------------ ac_fir_reg_share
@pytest.mark.parametrize("cid", range(len(rc.RS_CONFIGS)), ids=lambda i: f"rs{i}-{rc.RS_CONFIGS[i][8]}-{rc.RS_CONFIGS[i][0]}")
def test_reg_share_restatement_vs_reference_outputs(oracle, cid):
    """Oracle B against the committed outputs of the UNMODIFIED reference class (tests/golden/rs_outputs.npz)."""
    g = golden("rs_outputs.npz")
    N, fi, fo, fc, fa, mww, bs, bo, ft = rc.RS_CONFIGS[cid]
    f = oracle.RsB(fi, fo, fc, fa, N, mww, bs, bo, ft)
    assert f.ram_words == rc.rs_ram_words(rc.RS_CONFIGS[cid]) == g[f"rs{cid}_ram"].size
    x, ram = g[f"rs{cid}_x"], g[f"rs{cid}_ram"]
    y = np.concatenate([f.run(x[:7], ram), f.run(x[7:8], ram), f.run(x[8:], ram)])
    assert np.array_equal(y, g[f"rs{cid}_y"])
    assert f.delay_out() == int(g[f"rs{cid}_dl"][0])


def test_reg_share_live_reference_equals_restatement(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no reference tree)")
    rng = np.random.default_rng(3)
    for cid, (N, fi, fo, fc, fa, mww, bs, bo, ft) in enumerate(rc.RS_CONFIGS):
        a, b = oracle.RsA(cid), oracle.RsB(fi, fo, fc, fa, N, mww, bs, bo, ft)
        for kind in ("uniform", "min", "max", "alt"):
            x, ram = oracle.rand_raw(rng, fi, 3 * N + 5, kind), oracle.rand_raw(rng, fc, a.ram_words, "uniform" if kind == "alt" else kind)
            assert np.array_equal(a.run(x, ram), b.run(x, ram)), (cid, kind)
            assert a.delay_out() == b.delay_out()


def test_reg_share_kat_antisymmetric_kills_dc(oracle):
    """Derived from the semantics (not a reference test): an anti-symmetric fold of a constant input is zero."""
    f = oracle.RsB((16, 1), (40, 8), (16, 1), (40, 8), 16, 1, 1, 0, "FOLD_EVEN_ANTI")
    y = f.run(np.full(64, 1234), np.arange(1, 9) * 1000)
    assert np.all(y[15:] == 0) and np.any(y[:15] != 0)


# ------------------------------------------------------------------------ ac_poly_dec
@pytest.mark.parametrize("cid", range(len(rc.PD_CONFIGS)), ids=lambda i: f"pd{i}-NT{rc.PD_CONFIGS[i][4]}-DF{rc.PD_CONFIGS[i][5]}")
def test_poly_dec_restatement_vs_reference_outputs(oracle, cid):
    g = golden("rs_outputs.npz")
    fi, fc, fa, fo, nt, df = rc.PD_CONFIGS[cid]
    f = oracle.PdB(fi, fc, fa, fo, nt, df)
    f.load(g[f"pd{cid}_c"])
    x = g[f"pd{cid}_x"]
    y = np.concatenate([f.run(x[:1]), f.run(x[1:df + 2]), f.run(x[df + 2:])])
    assert y.size == x.size // df and np.array_equal(y, g[f"pd{cid}_y"])


def test_poly_dec_live_reference_equals_restatement(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no reference tree)")
    rng = np.random.default_rng(4)
    for cid, (fi, fc, fa, fo, nt, df) in enumerate(rc.PD_CONFIGS):
        a, b = oracle.PdA(cid), oracle.PdB(fi, fc, fa, fo, nt, df)
        for kind in ("uniform", "min", "max", "alt"):
            c = oracle.rand_raw(rng, fc, nt * df, "uniform" if kind == "alt" else kind)
            a.load(c); b.load(c)
            x = oracle.rand_raw(rng, fi, 5 * nt * df + 3, kind)
            assert np.array_equal(a.run(x), b.run(x)), (cid, kind)


# ------------------------------------------------------------------------ ac_poly_intr
def _pi_ids(i):
    return f"pi{i}-{rc.PI_CONFIGS[i][6]}-NT{rc.PI_CONFIGS[i][4]}-IF{rc.PI_CONFIGS[i][5]}"


def pi_replay(f, g, cid, cuts=(1, 9)):
    """The call sequence the fixture was generated with: first coefficient / control set, three runs, reload, one run."""
    x, half = g[f"pi{cid}_x"], int(g[f"pi{cid}_half"][0])
    sign, corr = g[f"pi{cid}_sign"], g[f"pi{cid}_corr"]
    f.load(g[f"pi{cid}_c1"], sign, corr)
    ys = [f.run(x[:cuts[0]]), f.run(x[cuts[0]:cuts[1]]), f.run(x[cuts[1]:half])]
    f.load(g[f"pi{cid}_c2"], 1 - sign, corr[::-1].copy())
    ys.append(f.run(x[half:]))
    return np.concatenate([np.asarray(y).reshape(-1) for y in ys])


@pytest.mark.parametrize("cid", range(len(rc.PI_CONFIGS)), ids=_pi_ids)
def test_poly_intr_restatement_vs_reference_outputs(oracle, cid):
    g = golden("rs_outputs.npz")
    fi, fc, fa, fo, nt, IF, ft = rc.PI_CONFIGS[cid]
    y = pi_replay(oracle.PiB(fi, fc, fa, fo, nt, IF, ft), g, cid)
    n = g[f"pi{cid}_x"].size
    assert y.size == IF * (n if ft == "FOLD_ANTI" else n - 1)        # the folded forms write one step late (ac_poly_intr.h:160)
    assert np.array_equal(y, g[f"pi{cid}_y"])


def test_poly_intr_live_reference_equals_restatement(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no reference tree)")
    rng = np.random.default_rng(41)
    for cid, (fi, fc, fa, fo, nt, IF, ft) in enumerate(rc.PI_CONFIGS):
        a, b = oracle.PiA(cid), oracle.PiB(fi, fc, fa, fo, nt, IF, ft)
        for kind in ("uniform", "min", "max", "alt"):
            c = oracle.rand_raw(rng, fc, a.coeffsz, "uniform" if kind == "alt" else kind)
            sign, corr = rng.integers(0, 2, size=IF), rng.integers(0, IF, size=IF)
            a.load(c, sign, corr); b.load(c, sign, corr)
            x = oracle.rand_raw(rng, fi, 4 * nt + 9, kind)
            assert np.array_equal(a.run(x), b.run(x)), (cid, kind)


def test_poly_intr_kat_plain_form_is_the_polyphase_fir(oracle):
    """Derived from the semantics (not a reference test): FOLD_ANTI with exact accumulators equals the zero-stuffed input
    filtered by the interleaved prototype h[IF*m + j] = coeffs[m + NTAPS*j]; a symmetric pair (corr = mirror phase) of the
    folded form returns (sum + difference) / 2 of the two parked accumulators."""
    rng = np.random.default_rng(8)
    NT, IF = 6, 4
    c = rng.integers(-2000, 2000, size=NT * IF)
    x = rng.integers(-3000, 3000, size=50)
    f = oracle.PiB((16, 1), (16, 1), (40, 8), (40, 8), NT, IF, "FOLD_ANTI")
    f.load(c)
    y = f.run(x)
    proto = np.zeros(NT * IF, dtype=np.int64)
    for j in range(IF):
        proto[j::IF] = c[j * NT:(j + 1) * NT]
    z = np.zeros(x.size * IF, dtype=np.int64)
    z[::IF] = x
    assert np.array_equal(y, np.convolve(z, proto)[:z.size] << 2)       # F_acc - F_in - F_c = 2
    e = oracle.PiB((16, 1), (16, 1), (40, 8), (40, 8), NT, 2, "FOLD_EVEN")
    ce = rng.integers(-2000, 2000, size=NT)
    e.load(ce, [1, 0], [1, 0])
    ye = e.run(x).reshape(-1, 2)
    plain = oracle.PiB((16, 1), (16, 1), (40, 8), (40, 8), NT, 2, "FOLD_EVEN")
    plain.load(ce, [1, 0], [0, 1])
    yp = plain.run(x).reshape(-1, 2)                                    # the parked accumulators themselves: S (sum set), D (difference set)
    assert np.array_equal(ye[:, 0], (yp[:, 0] - yp[:, 1]) >> 1) and np.array_equal(ye[:, 1], (yp[:, 1] + yp[:, 0]) >> 1)


# ------------------------------------------------------------------------ ac_intg_dump
def _id_calls(g, cid):
    x, ns, xlen = g[f"id{cid}_x"], g[f"id{cid}_ns"], g[f"id{cid}_xlen"]
    o = 0
    for call in range(3):
        yield x[o:o + xlen[call]], ns[6 * call:6 * call + 6]
        o += xlen[call]


@pytest.mark.parametrize("cid", range(len(rc.ID_CONFIGS)), ids=lambda i: f"id{i}-NS{rc.ID_CONFIGS[i][3]}-CHN{rc.ID_CONFIGS[i][4]}")
def test_intg_dump_restatement_vs_reference_outputs(oracle, cid):
    g = golden("rs_outputs.npz")
    fi, fa, fo, NS, CHN = rc.ID_CONFIGS[cid]
    f = oracle.IdB(fi, fa, fo, NS, CHN)
    y = np.concatenate([f.run(x, ns) for x, ns in _id_calls(g, cid)])
    assert np.array_equal(y, g[f"id{cid}_y"])


# ------------------------------------------------------------------------ ac_mv_avg (parity unpinned)
MV_IDS = [f"mv{i}-{c[2]}-{c[1]}" for i, c in enumerate(rc.MV_CONFIGS)]


@pytest.mark.parametrize("cid", range(len(rc.MV_CONFIGS)), ids=MV_IDS)
def test_mv_avg_restatement_vs_reference_outputs(oracle, cid):
    """Oracle B against the committed outputs of the UNMODIFIED ac_mv_avg.h driven over the restated window class."""
    g = golden("rs_outputs.npz")
    maxs, taps, wt, fi, fo, fa, fc = rc.MV_CONFIGS[cid]
    ns1, ns2 = (int(v) for v in g[f"mv{cid}_ns"])
    c = g[f"mv{cid}_c"]
    assert np.array_equal(oracle.mv_run_b(fi, fo, fa, fc, taps, wt, c, g[f"mv{cid}_x1"], ns1), g[f"mv{cid}_y1"])
    assert np.array_equal(oracle.mv_run_b(fi, fo, fa, fc, taps, wt, c, g[f"mv{cid}_x2"], ns2), g[f"mv{cid}_y2"])
    assert g[f"mv{cid}_y1"].size == 3 * (ns1 - taps + 1 if wt == "AC_WIN" else ns1)


def test_mv_avg_documented_behaviour(oracle):
    """The manual's description (section 2.4.3), on the restatement: DC gain = sum of the weights everywhere (clip and
    mirror keep a constant burst constant), an impulse in the middle returns the reversed weights, the two boundary
    rules differ only within TAPS/2 of the burst edges, AC_WIN is the interior of either."""
    Q = (16, 2)
    A = (32, 6)
    taps, n = 7, 50
    c = np.array([3, -1, 4, 1, -5, 9, 2]) * 256
    dc = np.full(n, 1 << 10)
    for wt in ("AC_CLIP", "AC_MIRROR"):
        y = oracle.mv_run_b(Q, A, A, Q, taps, wt, c, dc, n)
        assert y.size == n and np.all(y == ((dc[0] * c) >> 2).sum())      # (x * 2^12 * c) >> 14 per tap: floor(x c / 4)
    x = np.zeros(n, dtype=np.int64)
    x[25] = 1 << 14
    y = oracle.mv_run_b(Q, A, A, Q, taps, "AC_CLIP", c, x, n)
    assert np.array_equal(y[22:29], c[::-1] * (1 << 14) >> 2)
    rng = np.random.default_rng(8)
    x = rng.integers(-30000, 30000, n)
    yc, ym, yw = (oracle.mv_run_b(Q, A, A, Q, taps, wt, c, x, n) for wt in ("AC_CLIP", "AC_MIRROR", "AC_WIN"))
    assert np.array_equal(yc[3:-3], ym[3:-3]) and np.array_equal(yw, yc[3:-3]) and not np.array_equal(yc[:3], ym[:3])
    # the left edge by hand: clip repeats x[0], mirror reflects about it
    k = 1
    idx_c = [max(k + j, 0) for j in range(-3, 4)]
    idx_m = [abs(k + j) for j in range(-3, 4)]
    assert yc[k] == sum((int(x[i]) * int(w)) >> 2 for i, w in zip(idx_c, c)) and ym[k] == sum((int(x[i]) * int(w)) >> 2 for i, w in zip(idx_m, c))
    with pytest.raises(ValueError):
        oracle.mv_run_b(Q, A, A, Q, taps, "AC_CLIP", c, x[:6], 6)           # burst shorter than the window
