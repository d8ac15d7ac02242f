"""Oracle A == Oracle B on RANDOM instantiations (dev container only: needs /root/reference).

tests/test_oracle.py compares the two oracles on the hand-picked table in oracle/ref_configs.py.  Here the table is
drawn at random -- formats (widths, binary points, signedness), all 8 quantisation and 4 overflow modes for the
accumulator and the output, tap counts, R / M / N -- the unmodified reference templates are compiled for exactly
those instantiations (a throw-away .so next to the test's tmp dir, same driver sources as oracle/_ref), and the
plain-C restatement has to reproduce them bit for bit, single call and chunked.  Oracle B is what checks the CUDA
engine on the GPU box, so this is the widest net under the parity chain.  B2D_FUZZ_SEED picks another draw.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("AC_DSP_REF", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "include", "ac_dsp", "ac_fir_load_coeffs.h")),
                                reason="reference tree not present (GPU box): the committed fixtures pin Oracle B there")
SEED = int(os.environ.get("B2D_FUZZ_SEED", "20260101"))
Q_MODES, O_MODES = O.Q_MODES, O.O_MODES


def cfmt(f):
    W, I, S, Q, Om = f
    return f"{W},{I},{'true' if S else 'false'},{Q},{Om}"


def compile_driver(tmp, incs, sources, name):
    """incs: {file name: text} placed under tmp/_ref/.  The drivers #include "_ref/cfgs_*.inc" with quotes, which is
    looked up next to the including file first, so the (test-infrastructure) driver sources are copied into tmp and
    compiled there: these lists replace oracle/_ref's."""
    import shutil
    os.makedirs(os.path.join(tmp, "_ref"), exist_ok=True)
    for fn, text in incs.items():
        with open(os.path.join(tmp, "_ref", fn), "w") as fh:
            fh.write(text)
    shutil.copy(os.path.join(ROOT, "oracle", "ref_driver_cic.h"), tmp)
    objs = []
    for src, defs in sources:
        local = shutil.copy(os.path.join(ROOT, "oracle", src), tmp)
        obj = os.path.join(tmp, src + "".join(defs).replace("-D", "_") + ".o")
        subprocess.check_call(["g++", "-std=c++11", "-O1", "-fPIC", f"-I{tmp}", f"-I{ROOT}/oracle/ac_shim", f"-I{REF}/include",
                               "-c", local, "-o", obj] + defs)
        objs.append(obj)
    lib = os.path.join(tmp, name)
    subprocess.check_call(["g++", "-shared", "-o", lib] + objs)
    return C.CDLL(lib)


def rand_fmt(rng, wlo, whi, modes=False):
    W = int(rng.integers(wlo, whi + 1))
    I = int(rng.integers(-2, W + 3))
    S = bool(rng.integers(0, 2)) or W == 1
    Q = Q_MODES[int(rng.integers(0, 8))] if modes else "AC_TRN"
    Om = O_MODES[int(rng.integers(0, 4))] if modes else "AC_WRAP"
    return (W, I, S, Q, Om)


# ------------------------------------------------------------------------------------------------ FIR
def draw_fir(rng, k):
    cfgs = []
    while len(cfgs) < k:
        fi, fc = rand_fmt(rng, 2, 32), rand_fmt(rng, 2, 32)
        Fp = (fi[0] - fi[1]) + (fc[0] - fc[1])
        Wa = int(rng.integers(8, 65))
        Fa = Fp + int(rng.integers(-12, 7))                    # mostly dropping bits per tap, sometimes an exact shift
        fa = (Wa, Wa - Fa, True if rng.integers(0, 4) else False, Q_MODES[int(rng.integers(0, 8))], O_MODES[int(rng.integers(0, 4))])
        Wo = int(rng.integers(4, 65))
        fo = (Wo, Wo - (Fa - int(rng.integers(0, 10))), bool(rng.integers(0, 2)), Q_MODES[int(rng.integers(0, 8))], O_MODES[int(rng.integers(0, 4))])
        nt = int(rng.choice([2, 3, 4, 5, 8, 11, 16, 27, 40]))
        # what the 128-bit shim (and the engine's 128-bit generic path) can hold: product, fold product, sums
        if fi[0] + fc[0] > 64 or fc[0] + fa[0] > 96:
            continue
        if abs(fa[1]) > 70 or abs(fo[1]) > 70 or max(Fa, Fp) - min(Fa, Fp) > 40:
            continue
        cfgs.append((fi, fc, fa, fo, nt))
    return cfgs


@pytest.fixture(scope="module")
def fir_fuzz(tmp_path_factory):
    rng = np.random.default_rng(SEED)
    cfgs = draw_fir(rng, 10)
    inc = "".join(f"X({i}, {cfmt(fi)}, {cfmt(fc)}, {cfmt(fa)}, {cfmt(fo)}, {nt})\n" for i, (fi, fc, fa, fo, nt) in enumerate(cfgs))
    L = compile_driver(str(tmp_path_factory.mktemp("firfuzz")), {"cfgs_fir.inc": inc}, [("ref_driver_fir.cpp", [])], "libfirfuzz.so")
    L.acref_fir_create.restype = C.c_void_p
    L.acref_fir_create.argtypes = [C.c_int, C.c_int, C.c_int]
    L.acref_fir_load.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.acref_fir_run.restype = C.c_long
    L.acref_fir_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64)]
    L.acref_fir_destroy.argtypes = [C.c_void_p]
    return L, cfgs


def p64(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


@pytest.mark.parametrize("i", range(10))
def test_fir_random_instantiation(fir_fuzz, i):
    L, cfgs = fir_fuzz
    fi, fc, fa, fo, nt = cfgs[i]
    rng = np.random.default_rng(SEED + 100 + i)
    x = O.rand_raw(rng, fi, 260)
    x[:4] = [O.rand_raw(rng, fi, 1, "min")[0], O.rand_raw(rng, fi, 1, "max")[0], 0, O.rand_raw(rng, fi, 1, "min")[0]]
    h = O.rand_raw(rng, fc, nt)
    h2 = np.ascontiguousarray(O.rand_raw(rng, fc, nt), dtype=np.int64)
    for ft in ("SHIFT_REG", "ROTATE_SHIFT", "C_BUFF", "FOLD_EVEN", "FOLD_ODD", "TRANSPOSED"):
        # FOLD_EVEN on an odd count / FOLD_ODD on an even one are not the intended use but compile: compared as well
        for cls in (0, 1, 2):
            ha = L.acref_fir_create(i, cls, O.FTYPES.index(ft))
            assert ha, (cfgs[i], ft, cls)
            hh = np.ascontiguousarray(h, dtype=np.int64)
            assert L.acref_fir_load(ha, p64(hh)) == 0
            xa = np.ascontiguousarray(x, dtype=np.int64)
            ya = np.empty(x.size, dtype=np.int64)
            # the reference in two calls (state carried by the object), the restatement in three differently cut ones
            n1 = L.acref_fir_run(ha, p64(xa), 97, p64(ya))
            reload_ = cls != 0                             # load / prog classes: another coefficient set mid-stream, delay line kept
            if reload_:
                assert L.acref_fir_load(ha, p64(h2)) == 0
            rest = np.ascontiguousarray(xa[97:])
            yb_ = np.empty(rest.size, dtype=np.int64)
            n2 = L.acref_fir_run(ha, p64(rest), rest.size, p64(yb_))
            L.acref_fir_destroy(ha)
            assert n1 == 97 and n2 == rest.size
            ya = np.concatenate([ya[:97], yb_])
            b = O.FirB(fi, fc, fa, fo, nt, ft)
            b.load(h)
            yb = [b.run(x[:1]), b.run(x[1:97])]
            if reload_:
                b.load(h2)
            yb = np.concatenate(yb + [b.run(x[97:130]), b.run(x[130:])])
            assert np.array_equal(ya, yb), (cfgs[i], ft, cls, int(np.flatnonzero(ya != yb)[0]))


# ------------------------------------------------------------------------------------------------ CIC
def draw_cic(rng, k, mode):
    cfgs = []
    while len(cfgs) < k:
        R, M, N = int(rng.integers(2, 10)), int(rng.integers(1, 4)), int(rng.integers(1, 6))
        if len(cfgs) == 0:
            R, M, N = 256, 1, 2                                    # the 8-bit rate counter's limit (ac_cic_full_core.h:72-73)
        elif len(cfgs) == 1:
            R, M, N = int(rng.integers(100, 256)), 2, int(rng.integers(1, 3))
        elif len(cfgs) == 2:
            R, M, N = int(rng.integers(2, 5)), 1, int(rng.integers(6, 9))
        W = int(rng.integers(3, 25))
        fi = (W, int(rng.integers(-1, W + 2)), bool(rng.integers(0, 2)) or W < 4, "AC_TRN", "AC_WRAP")
        intW = O.cic_int_width(mode, fi, R, M, N)
        if intW > 64:
            continue
        Fin = fi[0] - fi[1]
        Wo = int(rng.integers(4, 65))
        fo = (Wo, Wo - (Fin - int(rng.integers(0, 6))), bool(rng.integers(0, 2)), Q_MODES[int(rng.integers(0, 8))], O_MODES[int(rng.integers(0, 4))])
        if rng.integers(0, 3) == 0:
            fo = (intW, intW - Fin, True, "AC_TRN", "AC_WRAP")      # the lossless type, passed on unchanged
        cfgs.append((R, M, N, fi, fo))
    return cfgs


@pytest.fixture(scope="module")
def cic_fuzz(tmp_path_factory):
    rng = np.random.default_rng(SEED + 7)
    dec, intr = draw_cic(rng, 8, "dec"), draw_cic(rng, 8, "intr")
    incs = {"cfgs_cic_dec.inc": "".join(f"X({i}, {R}, {M}, {N}, {cfmt(fi)}, {cfmt(fo)})\n" for i, (R, M, N, fi, fo) in enumerate(dec)),
            "cfgs_cic_intr.inc": "".join(f"X({i}, {R}, {M}, {N}, {cfmt(fi)}, {cfmt(fo)})\n" for i, (R, M, N, fi, fo) in enumerate(intr))}
    L = compile_driver(str(tmp_path_factory.mktemp("cicfuzz")), incs,
                       [("ref_driver_cic.cpp", ["-DACREF_CIC_DEC"]), ("ref_driver_cic.cpp", ["-DACREF_CIC_INTR"])], "libcicfuzz.so")
    for fn in (L.acref_cic_dec_create, L.acref_cic_intr_create):
        fn.restype = C.c_void_p
        fn.argtypes = [C.c_int]
    L.acref_cic_run.restype = C.c_long
    L.acref_cic_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_long, C.POINTER(C.c_int64)]
    L.acref_cic_destroy.argtypes = [C.c_void_p]
    return L, dec, intr


@pytest.mark.parametrize("mode", ["dec", "intr"])
@pytest.mark.parametrize("i", range(8))
def test_cic_random_instantiation(cic_fuzz, mode, i):
    L, dec, intr = cic_fuzz
    R, M, N, fi, fo = (dec if mode == "dec" else intr)[i]
    rng = np.random.default_rng(SEED + 300 + i)
    x = np.ascontiguousarray(O.rand_raw(rng, fi, 400 if mode == "intr" or R < 50 else 3 * R + 77), dtype=np.int64)
    ha = (L.acref_cic_dec_create if mode == "dec" else L.acref_cic_intr_create)(i)
    assert ha
    outs = []
    for lo, hi in ((0, 1), (1, 58), (58, 61), (61, x.size)):         # the reference itself in ragged calls
        seg = np.ascontiguousarray(x[lo:hi])
        buf = np.empty(seg.size * R + R + 8, dtype=np.int64)
        n = L.acref_cic_run(ha, p64(seg), seg.size, p64(buf))
        outs.append(buf[:n].copy())
    L.acref_cic_destroy(ha)
    ya = np.concatenate(outs)
    b = O.CicB(mode, fi, fo, R, M, N)
    yb = np.concatenate([b.run(x[:200]), b.run(x[200:])])
    assert ya.size == yb.size and np.array_equal(ya, yb), ((R, M, N, fi, fo), ya.size, yb.size)


# ------------------------------------------------------------------------------- ac_poly_dec / ac_poly_intr / ac_intg_dump
def draw_mac_formats(rng):
    """(in, coeff, acc, out) like draw_fir: random binary points, all modes on the accumulator and the output."""
    while True:
        fi, fc = rand_fmt(rng, 2, 32), rand_fmt(rng, 2, 32)
        Fp = (fi[0] - fi[1]) + (fc[0] - fc[1])
        Wa = int(rng.integers(8, 65))
        Fa = Fp + int(rng.integers(-12, 7))
        fa = (Wa, Wa - Fa, True if rng.integers(0, 4) else False, Q_MODES[int(rng.integers(0, 8))], O_MODES[int(rng.integers(0, 4))])
        Wo = int(rng.integers(4, 65))
        fo = (Wo, Wo - (Fa - int(rng.integers(0, 10))), bool(rng.integers(0, 2)), Q_MODES[int(rng.integers(0, 8))], O_MODES[int(rng.integers(0, 4))])
        if fi[0] + fc[0] > 64 or fc[0] + fa[0] > 96 or abs(fa[1]) > 70 or abs(fo[1]) > 70 or abs(Fa - Fp) > 40:
            continue
        return fi, fc, fa, fo


@pytest.fixture(scope="module")
def poly_fuzz(tmp_path_factory):
    rng = np.random.default_rng(SEED + 11)
    pd = [draw_mac_formats(rng) + (int(rng.integers(1, 13)), int(rng.integers(2, 7))) for _ in range(8)]
    pi = []
    for _ in range(9):
        ft = ["FOLD_EVEN", "FOLD_ODD", "FOLD_ANTI"][len(pi) % 3]
        nt = int(rng.integers(1, 7)) * 2 if ft == "FOLD_EVEN" else (int(rng.integers(0, 6)) * 2 + 1 if ft == "FOLD_ODD" else int(rng.integers(1, 12)))
        pi.append(draw_mac_formats(rng) + (nt, int(rng.integers(1, 6)), ft))
    idc = []
    for _ in range(8):
        fi = rand_fmt(rng, 2, 32)
        Fa = (fi[0] - fi[1]) + int(rng.integers(-8, 5))
        Wa = int(rng.integers(8, 65))
        fa = (Wa, Wa - Fa, bool(rng.integers(0, 4)), Q_MODES[int(rng.integers(0, 8))], O_MODES[int(rng.integers(0, 4))])
        Wo = int(rng.integers(4, 65))
        fo = (Wo, Wo - (Fa - int(rng.integers(0, 8))), bool(rng.integers(0, 2)), Q_MODES[int(rng.integers(0, 8))], O_MODES[int(rng.integers(0, 4))])
        idc.append((fi, fa, fo, int(rng.integers(4, 65)), int(rng.integers(1, 6))))
    from oracle import ref_configs as rc
    incs = {
        "cfgs_pd.inc": "".join(f"X({i}, {cfmt(a)}, {cfmt(b)}, {cfmt(c)}, {cfmt(d)}, {nt}, {df})\n" for i, (a, b, c, d, nt, df) in enumerate(pd)),
        "cfgs_pi.inc": "".join(f"X({i}, {cfmt(a)}, {cfmt(b)}, {cfmt(c)}, {cfmt(d)}, {nt}, {rc.pi_coeffsz((a, b, c, d, nt, IF, ft))}, {IF}, {ft})\n"
                               for i, (a, b, c, d, nt, IF, ft) in enumerate(pi)),
        "cfgs_id.inc": "".join(f"X({i}, {cfmt(a)}, {cfmt(b)}, {cfmt(c)}, {ns}, {chn})\n" for i, (a, b, c, ns, chn) in enumerate(idc)),
    }
    L = compile_driver(str(tmp_path_factory.mktemp("polyfuzz")), incs,
                       [("ref_driver_pd.cpp", []), ("ref_driver_pi.cpp", []), ("ref_driver_id.cpp", [])], "libpolyfuzz.so")
    P64 = C.POINTER(C.c_int64)
    for name in ("pd", "pi", "id"):
        getattr(L, f"acref_{name}_create").restype = C.c_void_p
        getattr(L, f"acref_{name}_create").argtypes = [C.c_int]
        getattr(L, f"acref_{name}_destroy").argtypes = [C.c_void_p]
    L.acref_pd_load.argtypes = [C.c_void_p, P64]
    L.acref_pd_run.restype = C.c_long
    L.acref_pd_run.argtypes = [C.c_void_p, P64, C.c_long, P64]
    L.acref_pi_load.argtypes = [C.c_void_p, P64, P64, P64]
    L.acref_pi_run.restype = C.c_long
    L.acref_pi_run.argtypes = [C.c_void_p, P64, C.c_long, P64]
    L.acref_id_run.restype = C.c_long
    L.acref_id_run.argtypes = [C.c_void_p, P64, C.c_long, P64, C.c_long, P64]
    return L, pd, pi, idc


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


@pytest.mark.parametrize("i", range(8))
def test_poly_dec_random_instantiation(poly_fuzz, i):
    L, pd, _pi, _id = poly_fuzz
    fi, fc, fa, fo, nt, df = pd[i]
    rng = np.random.default_rng(SEED + 500 + i)
    x, h = i64(O.rand_raw(rng, fi, 50 * df + 3)), i64(O.rand_raw(rng, fc, nt * df))
    h2 = i64(O.rand_raw(rng, fc, nt * df))
    ha = L.acref_pd_create(i)
    L.acref_pd_load(ha, p64(h))
    b = O.PdB(fi, fc, fa, fo, nt, df)
    b.load(h)
    outs, outs_b = [], []
    cut = 20 * df + 1                                  # a new coefficient set arrives with a group half consumed
    for lo, hi in ((0, 1), (1, df + 2), (df + 2, cut), (cut, x.size)):
        if lo == cut:
            L.acref_pd_load(ha, p64(h2))
            b.load(h2)
        seg = i64(x[lo:hi])
        buf = np.empty(seg.size // df + 3, dtype=np.int64)
        outs.append(buf[:L.acref_pd_run(ha, p64(seg), seg.size, p64(buf))].copy())
        outs_b.append(b.run(seg))
    L.acref_pd_destroy(ha)
    yb = np.concatenate(outs_b)
    ya = np.concatenate(outs)
    assert ya.size == yb.size and np.array_equal(ya, yb), (pd[i], ya.size, yb.size)


@pytest.mark.parametrize("i", range(9))
def test_poly_intr_random_instantiation(poly_fuzz, i):
    L, _pd, pi, _id = poly_fuzz
    fi, fc, fa, fo, nt, IF, ft = pi[i]
    from oracle import ref_configs as rc
    csz = rc.pi_coeffsz(pi[i])
    rng = np.random.default_rng(SEED + 600 + i)
    x, h = i64(O.rand_raw(rng, fi, 120)), i64(O.rand_raw(rng, fc, csz))
    x[:3] = [O.rand_raw(rng, fi, 1, "min")[0], O.rand_raw(rng, fi, 1, "max")[0], O.rand_raw(rng, fi, 1, "min")[0]]
    sign, corr = i64(rng.integers(0, 2, IF)), i64(rng.integers(0, IF, IF))
    h2, sign2, corr2 = i64(O.rand_raw(rng, fc, csz)), i64(rng.integers(0, 2, IF)), i64(rng.integers(0, IF, IF))
    ha = L.acref_pi_create(i)
    b = O.PiB(fi, fc, fa, fo, nt, IF, ft)
    ya, yb = [], []
    for (lo, hi), (c, s, r) in zip(((0, 70), (70, 120)), ((h, sign, corr), (h2, sign2, corr2))):     # a reload mid-stream
        L.acref_pi_load(ha, p64(c), p64(s), p64(r))
        seg = i64(x[lo:hi])
        buf = np.empty(seg.size * IF + 1, dtype=np.int64)
        ya.append(buf[:L.acref_pi_run(ha, p64(seg), seg.size, p64(buf))].copy())
        b.load(c, s, r)
        yb.append(b.run(seg))
    L.acref_pi_destroy(ha)
    ya, yb = np.concatenate(ya), np.concatenate(yb)
    assert ya.size == yb.size and np.array_equal(ya, yb), (pi[i], ya.size, yb.size)


@pytest.mark.parametrize("i", range(8))
def test_intg_dump_random_instantiation(poly_fuzz, i):
    L, _pd, _pi, idc = poly_fuzz
    fi, fa, fo, NS, CHN = idc[i]
    rng = np.random.default_rng(SEED + 700 + i)
    tok = rng.integers(1, NS + 1, 30)
    tok[[4, 11, 20]] = [0, NS + 5, NS]                # tokens outside 1..NS consume NS samples and dump nothing
    n = int(sum(O.id_frame_samples(v, NS, CHN) for v in tok))
    x = i64(O.rand_raw(rng, fi, n))
    ha = L.acref_id_create(i)
    ya = []
    cut_f = 13                                        # two calls, split at a frame boundary (sums of non-dumping frames run on)
    cut_s = int(sum(O.id_frame_samples(v, NS, CHN) for v in tok[:cut_f]))
    for xs, ts in ((x[:cut_s], tok[:cut_f]), (x[cut_s:], tok[cut_f:])):
        xs, ts = i64(xs), i64(ts)
        buf = np.empty(ts.size * CHN + 1, dtype=np.int64)
        ya.append(buf[:L.acref_id_run(ha, p64(xs), xs.size, p64(ts), ts.size, p64(buf))].copy())
    ya = np.concatenate(ya)
    L.acref_id_destroy(ha)
    yb = np.asarray(O.IdB(fi, fa, fo, NS, CHN).run(x, tok)).reshape(-1)
    assert ya.size == yb.size and np.array_equal(ya, yb), (idc[i], ya.size, yb.size)


# ------------------------------------------------------------------------------------------ ac_fir_reg_share
@pytest.fixture(scope="module")
def rs_fuzz(tmp_path_factory):
    from oracle import ref_configs as rc
    rng = np.random.default_rng(SEED + 13)
    cfgs = []
    while len(cfgs) < 10:
        ft = ["SHIFT_REG", "FOLD_EVEN", "FOLD_ODD", "FOLD_EVEN_ANTI", "FOLD_ODD_ANTI"][len(cfgs) % 5]
        N = int(rng.integers(1, 13)) * 2 if "EVEN" in ft else (int(rng.integers(0, 12)) * 2 + 1 if "ODD" in ft else int(rng.integers(1, 25)))
        used = N if ft == "SHIFT_REG" else (N // 2 if "EVEN" in ft else (N - 1) // 2 + 1)
        bs = int(rng.choice([d for d in range(1, used + 1) if used % d == 0]))     # the tap loop must be whole blocks (else the reference reads out of range)
        mww = bs + int(rng.integers(0, 4))
        bo = int(rng.integers(0, mww - bs + 1))
        fi, fc, fa, fo = draw_mac_formats(rng)
        cfgs.append((N, fi, fo, fc, fa, mww, bs, bo, ft))
    inc = "".join(f"X({i}, {N}, {cfmt(fi)}, {cfmt(fo)}, {cfmt(fc)}, {cfmt(fa)}, {mww}, {bs}, {bo}, {ft}, {rc.rs_ram_words(c)})\n"
                  for i, c in enumerate(cfgs) for (N, fi, fo, fc, fa, mww, bs, bo, ft) in [c])
    L = compile_driver(str(tmp_path_factory.mktemp("rsfuzz")), {"cfgs_rs.inc": inc}, [("ref_driver_rs.cpp", [])], "librsfuzz.so")
    P64 = C.POINTER(C.c_int64)
    L.acref_rs_create.restype = C.c_void_p
    L.acref_rs_create.argtypes = [C.c_int]
    L.acref_rs_run.restype = C.c_long
    L.acref_rs_run.argtypes = [C.c_void_p, P64, C.c_long, P64, P64]
    L.acref_rs_delay_out.restype = C.c_longlong
    L.acref_rs_delay_out.argtypes = [C.c_void_p]
    L.acref_rs_destroy.argtypes = [C.c_void_p]
    return L, cfgs


@pytest.mark.parametrize("i", range(10))
def test_reg_share_random_instantiation(rs_fuzz, i):
    from oracle import ref_configs as rc
    L, cfgs = rs_fuzz
    N, fi, fo, fc, fa, mww, bs, bo, ft = cfgs[i]
    rng = np.random.default_rng(SEED + 800 + i)
    x = i64(O.rand_raw(rng, fi, 4 * N + 30))
    ram = i64(O.rand_raw(rng, fc, rc.rs_ram_words(cfgs[i])))
    ram2 = i64(O.rand_raw(rng, fc, ram.size))                      # programmable: another coefficient set mid-stream
    ha = L.acref_rs_create(i)
    b = O.RsB(fi, fo, fc, fa, N, mww, bs, bo, ft)
    ya, yb = [], []
    for (lo, hi), r in (((0, 2 * N), ram), ((2 * N, x.size), ram2)):
        seg = i64(x[lo:hi])
        buf = np.empty(seg.size, dtype=np.int64)
        assert L.acref_rs_run(ha, p64(seg), seg.size, p64(r), p64(buf)) == seg.size
        ya.append(buf)
        yb.append(b.run(seg, r))
    assert np.array_equal(np.concatenate(ya), np.concatenate(yb)), cfgs[i]
    assert int(L.acref_rs_delay_out(ha)) == b.delay_out(), cfgs[i]
    L.acref_rs_destroy(ha)


# ------------------------------------------------------------------------------------------ ac_mv_avg (parity unpinned)
def draw_mv(rng, k):
    """(MAX_SAMPLE, TAPS, WIN_TYPE, in, out, acc, coeff): the product is ACC_TYPE x COEFF_TYPE here (the sample is cast to
    ACC_TYPE first), so the 128-bit budget bounds W_acc + W_coeff."""
    cfgs = []
    while len(cfgs) < k:
        fi, fc, fa, fo = draw_mac_formats(rng)
        if fa[0] + fc[0] > 96 or fa[0] + max(0, (fc[0] - fc[1])) > 110:
            continue
        taps = int(rng.integers(0, 8)) * 2 + 1
        cfgs.append((taps + int(rng.integers(0, 40)), taps, ["AC_WIN", "AC_CLIP", "AC_MIRROR"][len(cfgs) % 3], fi, fo, fa, fc))
    return cfgs


@pytest.fixture(scope="module")
def mv_fuzz(tmp_path_factory):
    rng = np.random.default_rng(SEED + 17)
    cfgs = draw_mv(rng, 9)
    inc = "".join(f"X({i}, {maxs}, {taps}, {wt}, {cfmt(fi)}, {cfmt(fo)}, {cfmt(fa)}, {cfmt(fc)})\n" for i, (maxs, taps, wt, fi, fo, fa, fc) in enumerate(cfgs))
    L = compile_driver(str(tmp_path_factory.mktemp("mvfuzz")), {"cfgs_mv.inc": inc}, [("ref_driver_mv.cpp", [])], "libmvfuzz.so")
    P64 = C.POINTER(C.c_int64)
    L.acref_mv_create.restype = C.c_void_p
    L.acref_mv_create.argtypes = [C.c_int, P64]
    L.acref_mv_run.restype = C.c_long
    L.acref_mv_run.argtypes = [C.c_void_p, P64, C.c_long, C.c_longlong, P64]
    L.acref_mv_destroy.argtypes = [C.c_void_p]
    return L, cfgs


@pytest.mark.parametrize("i", range(9))
def test_mv_avg_random_instantiation(mv_fuzz, i):
    """The unmodified ac_mv_avg.h (over the restated window class) == the plain-C restatement on drawn formats and modes:
    pins the arithmetic of the class -- the ACC_TYPE cast before the multiply, the per-tap re-quantisation, the burst loop --
    not the window's boundary rules, which both sides take from oracle/ac_shim/ac_window.h's reading of the manual."""
    L, cfgs = mv_fuzz
    maxs, taps, wt, fi, fo, fa, fc = cfgs[i]
    rng = np.random.default_rng(SEED + 900 + i)
    c = i64(O.rand_raw(rng, fc, taps))
    ha = L.acref_mv_create(i, p64(c))
    for ns in (taps, maxs):
        x = i64(O.rand_raw(rng, fi, 3 * ns))
        buf = np.empty(x.size + 1, dtype=np.int64)
        ya = buf[:L.acref_mv_run(ha, p64(x), x.size, ns, p64(buf))].copy()
        yb = O.mv_run_b(fi, fo, fa, fc, taps, wt, c, x, ns)
        assert ya.size == yb.size and np.array_equal(ya, yb), (cfgs[i], ns)
    L.acref_mv_destroy(ha)
