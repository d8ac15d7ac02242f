"""The C++ header facade (include/b200dsp/ac_dsp/*.h): the reference's class templates re-created on the C-ABI.

tests/cpp/facade_bench.cpp drives the facade the way the reference's rtest_*.cpp benches drive the reference
classes (ac_fixed values on ac_channel FIFOs); this file builds it, feeds it the committed fixtures and compares
bit-for-bit.  CPU part: it compiles (and, where the reference tree is present, the reference's UNMODIFIED benches
compile against the facade too) and fails loudly without a GPU.  GPU part: parity.
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
LIBDIR = os.path.join(ROOT, "ac_dsp_b200", "lib")
REF = os.environ.get("AC_DSP_REF", "/root/reference")
CXXFLAGS = ["-std=c++11", "-O1", "-Wall", "-Werror", f"-I{ROOT}/include/b200dsp", f"-I{ROOT}/oracle/ac_shim"]
LDFLAGS = [f"-L{LIBDIR}", "-lb200dsp", f"-Wl,-rpath,{LIBDIR}"]


@pytest.fixture(scope="module")
def bench_exe(engine, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("facade") / "facade_bench")
    subprocess.check_call(["g++"] + CXXFLAGS + [os.path.join(ROOT, "tests", "cpp", "facade_bench.cpp")] + LDFLAGS + ["-o", exe])
    return exe


def run_case(exe, tmp_path, case, x, coeffs=None, chunk=0):
    fin, fco, fout = (str(tmp_path / n) for n in ("in.txt", "coef.txt", "out.txt"))
    np.savetxt(fin, np.asarray(x, dtype=np.int64), fmt="%d")
    if coeffs is not None:
        np.savetxt(fco, np.asarray(coeffs, dtype=np.int64), fmt="%d")
    p = subprocess.run([exe, case, fin, fco if coeffs is not None else "-", fout, str(chunk)], capture_output=True, text=True)
    return p, (np.loadtxt(fout, dtype=np.int64, ndmin=1) if p.returncode == 0 else None)


def test_facade_compiles_and_fails_loudly_without_gpu(engine, bench_exe, tmp_path):
    if engine.load().b2d_device_count() > 0:
        pytest.skip("GPU present: covered by the parity tests")
    p, y = run_case(bench_exe, tmp_path, "q15_cic_dec", np.arange(64))
    assert p.returncode == 70 and "engine_error" in p.stderr, (p.returncode, p.stderr)


FACADE_HEADERS = ["ac_fir_const_coeffs", "ac_fir_load_coeffs", "ac_fir_prog_coeffs", "ac_cic_dec_full", "ac_cic_intr_full",
                  "ac_fir_reg_share", "ac_poly_dec", "ac_poly_intr", "ac_intg_dump"]


def test_facade_headers_are_self_contained_and_combinable(tmp_path):
    """Each facade header compiles on its own (and twice: include guards); all of them share one translation unit --
    including both CIC headers, which the reference cannot (unguarded `power` template, ac_cic_dec_full.h:94-108 /
    ac_cic_intr_full.h:90-98) -- except ac_poly_intr, whose polyphase FTYPE enum clashes with the FIR one in the
    reference as well."""
    def syntax_only(name, includes):
        src = tmp_path / (name + ".cpp")
        src.write_text("".join(f"#include <ac_dsp/{h}.h>\n" for h in includes) + "int main() { return 0; }\n")
        p = subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", f"-I{ROOT}/include/b200dsp",
                            f"-I{ROOT}/oracle/ac_shim", str(src)], capture_output=True, text=True)
        assert p.returncode == 0, (name, p.stderr[-2000:])
    for h in FACADE_HEADERS:
        syntax_only(h, [h, h])
    syntax_only("all", [h for h in FACADE_HEADERS if h != "ac_poly_intr"])


def test_facade_marshaling_round_trips_every_storage_tier(engine, tmp_path):
    """include/b200dsp/marshal.h moves ac_fixed values through slc / set_slc only; check it against the shim's
    canonical raw value for 16..64-bit signed and unsigned types (no engine call, CPU)."""
    exe = str(tmp_path / "marshal_roundtrip")
    subprocess.check_call(["g++"] + CXXFLAGS + [os.path.join(ROOT, "tests", "cpp", "marshal_roundtrip.cpp")] + LDFLAGS + ["-o", exe])
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0 and "bad=0" in p.stdout, p.stdout + p.stderr


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "tests", "rtest_ac_fir_load_coeffs.cpp")),
                    reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", ["ac_fir_const_coeffs", "ac_fir_load_coeffs", "ac_fir_prog_coeffs", "ac_cic_dec_full", "ac_cic_intr_full"])
def test_unmodified_reference_benches_compile_against_the_facade(engine, name, tmp_path):
    """Drop-in at the source level: the reference's own bench, untouched, with the facade directory in front of the
    reference's include path (its <ac_dsp/...> includes then resolve to the facade)."""
    obj = str(tmp_path / (name + ".o"))
    cmd = ["g++", "-std=c++11", "-O0", f"-I{ROOT}/include/b200dsp", f"-I{ROOT}/oracle/ac_shim", f"-I{REF}/include",
           f"-I{REF}/tests", "-c", os.path.join(REF, "tests", f"rtest_{name}.cpp"), "-o", obj, "-H"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    used = [l for l in p.stderr.splitlines() if l.startswith(".") and f"ac_dsp/{name}.h" in l]
    assert used and all("include/b200dsp/ac_dsp" in l for l in used), used
    # and it links against the engine
    subprocess.check_call(["g++", obj] + LDFLAGS + ["-o", str(tmp_path / name)])


FIR_BENCHES = {"bench_const": "fir_bench_const.npz", "bench_load": "fir_bench_load.npz", "bench_prog": "fir_bench_prog.npz"}


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(FIR_BENCHES))
@pytest.mark.parametrize("chunk", [0, 100])
def test_facade_fir_reference_benches(bench_exe, tmp_path, case, chunk):
    """Stimulus / taps of the reference's FIR benches; expected = the UNMODIFIED reference class's outputs (fixture)."""
    g = np.load(os.path.join(GOLDEN, FIR_BENCHES[case]))
    p, y = run_case(bench_exe, tmp_path, case, g["x"], g["coeffs"], chunk)
    assert p.returncode == 0, p.stderr
    assert np.array_equal(y, g["y"])
    # the bench's own pass criterion: SQNR against the MATLAB double reference >= 60 dB, and equal to the reference's
    F = int(g["fout"][0] - g["fout"][1])
    ref = g["ref_double"]
    if case == "bench_const":   # rtest_ac_fir_const_coeffs.cpp:174 casts the reference through OUT_TYPE
        ref = np.floor(ref * 2.0 ** F) / 2.0 ** F
    d = y.astype(np.float64) / 2.0 ** F
    sqnr = 10 * np.log10(np.sum(ref ** 2) / np.sum((d - ref) ** 2))
    assert sqnr >= 60 and abs(sqnr - float(g["sqnr"])) < 1e-6, (sqnr, float(g["sqnr"]))


@pytest.mark.gpu
@pytest.mark.parametrize("case,fx", [("bench_cic_dec", "cic_dec_golden.npz"), ("bench_cic_intr", "cic_intr_golden.npz")])
@pytest.mark.parametrize("chunk", [0, 333])
def test_facade_cic_golden_vectors(bench_exe, tmp_path, case, fx, chunk):
    """The reference's MATLAB bit-exact CIC vectors through the facade classes."""
    g = np.load(os.path.join(GOLDEN, fx))
    p, y = run_case(bench_exe, tmp_path, case, g["x"], None, chunk)
    assert p.returncode == 0, p.stderr
    n = len(g["ref"])
    assert len(y) >= n and np.array_equal(y[:n], g["ref"])


@pytest.mark.gpu
def test_facade_baseline_configs(bench_exe, tmp_path, oracle):
    rng = np.random.default_rng(77)
    q15, acc40 = (16, 1), (40, 8)
    for case, taps, ft, n, chunk in [("q15_const16", 16, "SHIFT_REG", 3000, 0), ("q15_load256", 256, "SHIFT_REG", 5000, 777),
                                     ("q15_prog1024", 1024, "TRANSPOSED", 40, 1), ("q15_prog1024", 1024, "TRANSPOSED", 3000, 0)]:
        x = oracle.rand_raw(rng, q15, n)
        h = oracle.rand_raw(rng, q15, taps)
        p, y = run_case(bench_exe, tmp_path, case, x, h, chunk)
        assert p.returncode == 0, (case, p.stderr)
        ob = oracle.FirB(q15, q15, acc40, acc40, taps, ft)
        ob.load(h)
        assert np.array_equal(y, ob.run(x)), case
    x = oracle.rand_raw(rng, q15, 4001)
    for case, mode, fout, R, M, N in [("q15_cic_dec", "dec", (28, 13), 8, 1, 4), ("q15_cic_intr", "intr", (20, 5), 4, 1, 3)]:
        for chunk in (0, 13):
            p, y = run_case(bench_exe, tmp_path, case, x, None, chunk)
            assert p.returncode == 0, (case, p.stderr)
            assert np.array_equal(y, oracle.CicB(mode, q15, fout, R, M, N).run(x)), (case, chunk)


@pytest.mark.gpu
@pytest.mark.parametrize("cid", [2, 4, 7, 9])
def test_facade_reg_share(bench_exe, tmp_path, cid):
    """ac_fir_reg_share through its facade class (caller-owned delay line, scalar run(), blocked coefficient RAM)
    against the committed outputs of the unmodified reference class."""
    g = np.load(os.path.join(GOLDEN, "rs_outputs.npz"))
    x = g[f"rs{cid}_x"][:60]
    p, y = run_case(bench_exe, tmp_path, f"rs{cid}", x, g[f"rs{cid}_ram"])
    assert p.returncode == 0, p.stderr
    assert np.array_equal(y[:-1], g[f"rs{cid}_y"][:60])
    from oracle import ref_configs as rc
    N, fi, fo = rc.RS_CONFIGS[cid][:3]
    assert y[-1] == int(x[60 - N]) << ((fo[0] - fo[1]) - (fi[0] - fi[1]))   # OUT_TYPE(reg[N_TAPS-1]): widening or equal formats here


@pytest.mark.gpu
@pytest.mark.parametrize("cid", [2, 6])
@pytest.mark.parametrize("chunk", [0, 13])
def test_facade_poly_dec(bench_exe, tmp_path, cid, chunk):
    """ac_poly_dec through its facade class (coefficient struct on a channel, groups of DF samples) against the
    committed outputs of the unmodified reference class."""
    g = np.load(os.path.join(GOLDEN, "rs_outputs.npz"))
    p, y = run_case(bench_exe, tmp_path, f"pd{cid}", g[f"pd{cid}_x"], g[f"pd{cid}_c"], chunk)
    assert p.returncode == 0, p.stderr
    assert np.array_equal(y, g[f"pd{cid}_y"])


@pytest.fixture(scope="module")
def poly_intr_exe(engine, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("facade_pi") / "facade_poly_intr")
    subprocess.check_call(["g++"] + CXXFLAGS + [os.path.join(ROOT, "tests", "cpp", "facade_poly_intr.cpp")] + LDFLAGS + ["-o", exe])
    return exe


def test_facade_poly_intr_compiles_and_fails_loudly_without_gpu(engine, poly_intr_exe, tmp_path):
    if engine.load().b2d_device_count() > 0:
        pytest.skip("GPU present: covered by the parity test")
    g = np.load(os.path.join(GOLDEN, "rs_outputs.npz"))
    ctl = np.concatenate([g["pi0_c1"], g["pi0_c2"], g["pi0_sign"], g["pi0_corr"], g["pi0_half"][:1]])
    p, y = run_case(poly_intr_exe, tmp_path, "pi0", g["pi0_x"], ctl)
    assert p.returncode == 70 and "engine_error" in p.stderr, (p.returncode, p.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("cid", [0, 1, 2, 9])
def test_facade_poly_intr(poly_intr_exe, tmp_path, cid):
    """ac_poly_intr through its facade class (one run() per read_ctrl token, control / coefficient structs on channels,
    a reload half way) against the committed outputs of the unmodified reference class."""
    g = np.load(os.path.join(GOLDEN, "rs_outputs.npz"))
    ctl = np.concatenate([g[f"pi{cid}_c1"], g[f"pi{cid}_c2"], g[f"pi{cid}_sign"], g[f"pi{cid}_corr"], g[f"pi{cid}_half"][:1]])
    p, y = run_case(poly_intr_exe, tmp_path, f"pi{cid}", g[f"pi{cid}_x"], ctl)
    assert p.returncode == 0, p.stderr
    assert np.array_equal(y, g[f"pi{cid}_y"])


# ------------------------------------------------------------------------ the reference's own benches, unmodified (row N3)
REF_BIN = os.path.join(ROOT, "oracle", "_ref")


def _exact_decimal(raw, F):
    """raw * 2^-F as an exact decimal string (what MATLAB's dlmwrite put into the reference's vector files)."""
    import decimal
    with decimal.localcontext() as ctx:
        ctx.prec = 120
        return format(decimal.Decimal(int(raw)) / (decimal.Decimal(2) ** F), "f")


def _write_vectors(d, name):
    """Regenerate the text vectors a reference bench opens in its working directory from the committed fixtures
    (the reference tree itself does not travel to the GPU box)."""
    if name.startswith("ac_fir"):
        cls = name.split("_")[2]
        g = np.load(os.path.join(GOLDEN, f"fir_bench_{cls}.npz"))
        with open(os.path.join(d, f"{name}_ref.txt"), "w") as f:
            f.write("\n".join(repr(float(v)) for v in g["ref_double"]) + "\n")
        Fc = int(g["fcoeff"][0] - g["fcoeff"][1])
        with open(os.path.join(d, f"{name}_cfg.txt"), "w") as f:      # whitespace-separated doubles (load / prog read it at run time)
            f.write("\n".join(_exact_decimal(v, Fc) for v in g["coeffs"]) + "\n")
        return float(g["sqnr"])
    if name == "ac_cic_dec_full":
        g = np.load(os.path.join(GOLDEN, "cic_dec_golden.npz"))
        x, ref, skip = g["x"][1:], g["ref"], 0          # the bench prepends the leading zero itself (rtest_ac_cic_dec_full.cpp:78-85)
    else:
        g = np.load(os.path.join(GOLDEN, "cic_intr_golden.npz"))
        x, ref, skip = g["x"], g["ref"], int(g["N"])    # the bench throws the first N reference values away (rtest_ac_cic_intr_full.cpp:99)
    Fi, Fo = int(g["fin"][0] - g["fin"][1]), int(g["fout"][0] - g["fout"][1])
    with open(os.path.join(d, f"{name}_input.txt"), "w") as f:
        f.write("\n".join(_exact_decimal(v, Fi) for v in x) + "\n")
    with open(os.path.join(d, f"{name}_ref.txt"), "w") as f:
        f.write("\n".join(["0"] * skip + [_exact_decimal(v, Fo) for v in ref]) + "\n")
    return None


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ac_fir_const_coeffs", "ac_fir_load_coeffs", "ac_fir_prog_coeffs", "ac_cic_dec_full", "ac_cic_intr_full"])
def test_unmodified_reference_bench_passes_on_the_engine(name, tmp_path):
    """The reference's own rtest_<name>.cpp -- compiled UNMODIFIED in the dev container against the header facade and
    linked with libb200dsp.so (oracle/Makefile: facade_rtest_*) -- runs on the GPU and reports what it reports for the
    reference implementation: bit-exact CIC vectors, FIR SQNR 84.2385 / 89.5576 / 89.5576 dB."""
    exe = os.path.join(REF_BIN, f"facade_rtest_{name}")
    if not os.path.exists(exe):
        pytest.skip("facade_rtest binaries are built where the reference tree is present (oracle/Makefile)")
    want_sqnr = _write_vectors(str(tmp_path), name)
    p = subprocess.run([exe], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-2000:]
    assert "PASSED" in out and "FAILED" not in out, out[-2000:]
    if want_sqnr is not None:
        import re
        m = re.search(r"SQNR = ([0-9.]+)dB", out)
        assert m and abs(float(m.group(1)) - want_sqnr) < 1e-3, out[-500:]
    else:
        assert "Data mismatch" not in out


@pytest.mark.gpu
@pytest.mark.parametrize("cid", [0, 3])
def test_facade_intg_dump(bench_exe, tmp_path, cid):
    """ac_intg_dump through its facade class (token channel, frames split over several run() calls)."""
    g = np.load(os.path.join(GOLDEN, "rs_outputs.npz"))
    p, y = run_case(bench_exe, tmp_path, f"id{cid}", g[f"id{cid}_x"], g[f"id{cid}_ns"])
    assert p.returncode == 0, p.stderr
    assert np.array_equal(y, g[f"id{cid}_y"])


@pytest.mark.gpu
@pytest.mark.parametrize("cid", [0, 3, 8])
def test_facade_mv_avg(bench_exe, tmp_path, cid):
    """ac_mv_avg through its facade class (wrapper idiom, n_sample token channel): equal to the unmodified reference
    header driven over the restated window class (parity unpinned, oracle/ac_shim/ac_window.h)."""
    g = np.load(os.path.join(GOLDEN, "rs_outputs.npz"))
    for k in (1, 2):
        p, y = run_case(bench_exe, tmp_path, f"mv{cid}", g[f"mv{cid}_x{k}"], g[f"mv{cid}_c"], chunk=int(g[f"mv{cid}_ns"][k - 1]))
        assert p.returncode == 0, p.stderr
        assert np.array_equal(y, g[f"mv{cid}_y{k}"])
