"""Build recipe of libb200dsp.so: explicit nvcc for sm_100a, in-tree output (ac_dsp_b200/lib/).

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  There is no fallback:
if the library is missing and cannot be built, importing the engine fails.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libb200dsp.so")
SOURCES = ["runtime.cu", "rt_fir.cu", "rt_cic.cu", "rt_poly.cu", "rt_intgdump.cu", "rt_mvavg.cu", "mv_avg.cu", "wire.cu", "fir_generic.cu", "fir_q15.cu", "fir_ovs.cu", "fir_q24.cu", "fir_wide.cu", "fir_dec.cu", "fir_intr.cu", "intg_dump.cu", "upfir_q15.cu", "cic_generic.cu", "cic_fast.cu", "cic_intr_fast.cu", "nccl_dl.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b200dsp.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source into ac_dsp_b200/lib/libb200dsp.so (no-op when up to date)."""
    if not force and not stale():
        return LIB
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libb200dsp.so")
    os.makedirs(LIBDIR, exist_ok=True)
    tmp = LIB + ".tmp"
    # one nvcc -c per source in parallel, then one link
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])

    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [os.path.join(HERE, "..", "include", "b200dsp.h")]
    t_hdr = max(os.path.getmtime(f) for f in hdrs)

    def compile_one(src):
        obj = os.path.join(objdir, src + ".o")
        if not force and not verbose and os.path.exists(obj) and os.path.getmtime(obj) > max(t_hdr, os.path.getmtime(os.path.join(CSRC, src))):
            return obj                       # object newer than its source and every header: keep it
        subprocess.check_call([nvcc] + cflags + ["-c", os.path.join(CSRC, src), "-o", obj])
        return obj
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", tmp] + objs + ["-ldl"])
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
