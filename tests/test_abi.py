"""CPU suite: the C-ABI library loads, exports every symbol include/b200dsp.h declares, and its host-side
logic (descriptor validation, width calculators, sharding) behaves -- no compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    src = open(os.path.join(ROOT, "include", "b200dsp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2d_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(engine):
    lib = engine.load()
    names = declared_functions()
    assert len(names) >= 34
    for n in names:
        assert hasattr(lib, n), f"libb200dsp.so does not export {n}"


def test_container_bytes_and_strerror(engine):
    lib = engine.load()
    assert [lib.b2d_container_bytes(w) for w in (1, 16, 17, 32, 33, 64, 65, 0)] == [2, 2, 4, 4, 8, 8, 0, 0]
    assert lib.b2d_strerror(0) == b"ok" and b"supported" in lib.b2d_strerror(-1)
    assert b"sm_100a" in lib.b2d_version()


def test_int_width_matches_reference_formulas(engine):
    from ac_dsp_b200 import _lib as L
    lib = engine.load()
    cases = [(0, (16, 1), 8, 1, 4, 28), (0, (16, 1), 8, 2, 4, 32), (1, (16, 1), 4, 1, 3, 20),
             (0, (32, 16), 7, 2, 4, 48), (1, (32, 16), 7, 2, 5, 49), (0, (10, 2, False), 5, 1, 3, 18)]
    for mode, fin, R, M, N, want in cases:
        d = L.B2dCicDesc(L.make_fmt(fin), L.make_fmt((want, 1)), R, M, N, mode, 1, 0, 0)
        w = C.c_int32(0)
        assert lib.b2d_cic_int_width(C.byref(d), C.byref(w)) == 0
        assert w.value == want


def test_descriptor_validation_without_gpu(engine):
    """Rejections happen before any CUDA call, so they are observable on a CPU-only box."""
    from ac_dsp_b200 import _lib as L
    lib = engine.load()
    q15, acc = L.make_fmt((16, 1)), L.make_fmt((40, 8))
    h = C.c_void_p()

    def fir(**kw):
        base = dict(fin=q15, coeff=q15, acc=acc, out=acc, n_taps=16, ftype=0, kind=1, n_channels=1, layout=0, device=0)
        base.update(kw)
        return lib.b2d_fir_create(C.byref(h), C.byref(L.B2dFirDesc(*[base[k] for k, _ in L.B2dFirDesc._fields_])))

    assert fir(ftype=6) == L.EUNSUPPORTED and fir(ftype=7) == L.EUNSUPPORTED   # _ANTI: reference leaves output unwritten
    assert b"_ANTI" in lib.b2d_last_error()
    assert fir(n_taps=0) == L.EINVAL
    assert fir(n_channels=0) == L.EINVAL
    assert fir(fin=L.make_fmt((40, 8))) == L.EUNSUPPORTED                      # inputs wider than 32 bits
    assert fir(acc=L.make_fmt((65, 8))) == L.EINVAL
    assert fir(ftype=9) == L.EINVAL

    def cic(**kw):
        base = dict(fin=q15, out=L.make_fmt((28, 13)), R=8, M=1, N=4, mode=0, n_channels=1, layout=0, device=0)
        base.update(kw)
        return lib.b2d_cic_create(C.byref(h), C.byref(L.B2dCicDesc(*[base[k] for k, _ in L.B2dCicDesc._fields_])))

    assert cic(R=0) == L.EINVAL and cic(R=257) == L.EINVAL                     # 8-bit rate counters
    assert cic(R=1) == L.EUNSUPPORTED                                          # a valid instantiation the engine does not build
    assert cic(N=0) == L.EINVAL and cic(M=0) == L.EINVAL
    assert cic(R=256, N=8) == L.EUNSUPPORTED                                   # lossless width 80 > 64
    assert lib.b2d_device_count() >= 0


def test_wire_format_helpers(engine):
    """b2d_wire_bytes / b2d_unpack_wire are host-side: packed ceil(W/8)-byte values widen back to containers."""
    import numpy as np
    lib = engine.load()
    assert [lib.b2d_wire_bytes(w, 1) for w in (1, 8, 9, 16, 20, 24, 28, 32, 33, 40, 48, 56, 57, 64)] == [1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 8]
    assert [lib.b2d_wire_bytes(w, 0) for w in (8, 20, 40)] == [2, 4, 8] and lib.b2d_wire_bytes(40, 2) == 0
    rng = np.random.default_rng(4)
    for W, S in ((40, 1), (40, 0), (20, 1), (7, 1), (56, 1), (33, 0)):
        lo, hi = (-(1 << (W - 1)), (1 << (W - 1)) - 1) if S else (0, (1 << W) - 1)
        v = rng.integers(lo, hi, size=1001, endpoint=True, dtype=np.int64)
        v[:2] = [lo, hi]
        pb = (W + 7) // 8
        packed = np.zeros((v.size, pb), dtype=np.uint8)
        for b in range(pb):
            packed[:, b] = ((v.astype(np.uint64) >> np.uint64(8 * b)) & np.uint64(0xFF)).astype(np.uint8)
        out = np.zeros(v.size, dtype={2: np.int16, 4: np.int32, 8: np.int64}[lib.b2d_container_bytes(W)])
        assert lib.b2d_unpack_wire(packed.ctypes.data, v.size, W, S, out.ctypes.data) == 0
        assert np.array_equal(out.astype(np.int64), v), (W, S)


def test_shard_count(engine):
    lib = engine.load()
    n = C.c_uint32(0)
    for C_, world in ((64, 8), (8, 8), (10, 4), (3, 8), (1, 1)):
        tot = 0
        for r in range(world):
            assert lib.b2d_shard_count(C_, r, world, C.byref(n)) == 0
            assert n.value == len(range(r, C_, world))
            tot += n.value
        assert tot == C_
    assert lib.b2d_shard_count(8, 8, 8, C.byref(n)) != 0


def test_no_cpu_fallback_without_gpu(engine):
    """On a box without a GPU the engine must fail loudly rather than compute on the host."""
    if engine.load().b2d_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(engine.B2dError) as e:
        engine.ac_fir_load_coeffs((16, 1), (40, 8), (16, 1), (40, 8), 16)
    assert e.value.status == -3


def test_product_does_not_import_oracle():
    import subprocess
    import sys
    code = "import sys; import ac_dsp_b200; assert not any(m.startswith('oracle') for m in sys.modules), 'oracle imported'"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for dirpath, _d, files in os.walk(os.path.join(ROOT, "ac_dsp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("the CPU oracle", ""), f


def test_host_logic_fixed_point_core_and_packers(engine, tmp_path):
    """tests/cpp/host_logic_check.cu (host code only): csrc/common.cuh convert / macc / tap_term against the oracle
    shim's assignment and `+=` for every Q x O mode, and the coefficient packers against an independent expansion."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    libdir = os.path.join(ROOT, "ac_dsp_b200", "lib")
    exe = str(tmp_path / "host_logic_check")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-w", f"-I{ROOT}/ac_dsp_b200/csrc", f"-I{ROOT}/oracle/ac_shim",
                           os.path.join(ROOT, "tests", "cpp", "host_logic_check.cu"), "-o", exe,
                           f"-L{libdir}", "-lb200dsp", "-Xlinker", "-rpath", "-Xlinker", libdir])
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0 and "bad=0" in p.stdout, p.stdout[-3000:] + p.stderr[-1000:]
