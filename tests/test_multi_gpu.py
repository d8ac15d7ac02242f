"""N > 1 on real GPUs (-m gpu, skipped unless the box shows >= 2 devices): BASELINE configs[3] and configs[4] in small --
independent channels sharded over the ranks (channel c on rank c % world), ONE ncclBroadcast of the coefficient set
from rank 0 inside load(), every rank's outputs bit-equal to the oracle's for its channels, and the concatenation
over ranks identical to what a single GPU produces."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, zlib
sys.path.insert(0, os.environ["B2D_ROOT"])
import numpy as np
import torch
import torch.distributed as dist
import ac_dsp_b200 as E
from ac_dsp_b200 import parallel as P
from oracle import oracle as O
rank, world, local = P.rank_world()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = P.make_comm(rank, world, local)
Q15, ACC40 = (16, 1), (40, 8)
rng = np.random.default_rng(2026)                     # same stream on every rank: the job's full input
C, n, taps = 16, 4000, 1024          # >= 2 channels per rank at world = 8
x = rng.integers(-32768, 32767, size=(C, n), endpoint=True).astype(np.int16)
h = rng.integers(-32768, 32767, size=taps, endpoint=True).astype(np.int16)
mine = P.local_channels(C, rank, world)
# configs[3]: ac_fir_prog_coeffs, 1024 taps, channels sharded, taps known to rank 0 only
f = E.ac_fir_prog_coeffs(Q15, ACC40, Q15, ACC40, taps, "SHIFT_REG", n_channels=len(mine), layout="planar", device=local, comm=comm, root=0)
f.load(h if rank == 0 else None)
y = np.asarray(f.run(torch.from_numpy(x[mine]).cuda()).cpu().numpy()).reshape(len(mine), -1)
crc = {}
for i, c in enumerate(mine):
    ob = O.FirB(Q15, Q15, ACC40, ACC40, taps, "SHIFT_REG")
    ob.load(h)
    assert np.array_equal(y[i].astype(np.int64), ob.run(x[c])), ("fir", rank, c)
    crc[c] = zlib.crc32(y[i].astype(np.int64).tobytes())
# configs[4]: interpolator + 63-tap FIR cascade; its taps travel with the rendezvous (the cascade handle has no communicator)
g = rng.integers(-32768, 32767, size=63, endpoint=True).astype(np.int16)
cf = E.cic_intr_fir_cascade(Q15, (20, 5), 4, 1, 3, ACC40, Q15, ACC40, 63, "SHIFT_REG", n_channels=len(mine), device=local)
box = [g if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
cf.load(box[0])
yc = np.asarray(cf.run(torch.from_numpy(x[mine][:, :2000].copy()).cuda()).cpu().numpy()).reshape(len(mine), -1)
for i, c in enumerate(mine):
    oc, of = O.CicB("intr", Q15, (20, 5), 4, 1, 3), O.FirB((20, 5), Q15, ACC40, ACC40, 63, "SHIFT_REG")
    of.load(g)
    assert np.array_equal(yc[i].astype(np.int64), of.run(oc.run(x[c][:2000]))), ("cicfir", rank, c)
allc = [None] * world
dist.all_gather_object(allc, crc)
merged = {}
for d in allc:
    merged.update(d)
assert sorted(merged) == list(range(C))
if rank == 0:
    sys.stdout.write("CRC " + " ".join(str(merged[c]) for c in range(C)) + "\n")
dist.barrier()
f.close(); cf.close()
if comm: comm.close()
dist.destroy_process_group()
sys.stdout.write("rank%d-ok\n" % rank)
sys.stdout.flush()
'''


def _run(world, tmp_path):
    import socket
    w = tmp_path / f"worker{world}.py"
    w.write_text(WORKER)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    e = dict(os.environ, B2D_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(w)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=e, cwd=ROOT)
    if p.returncode != 0 or p.stdout.count("-ok") != world:      # keep the ranks' own words where a gpurun call brings them back
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", f"multi_gpu_world{world}_failure.txt"), "w") as fh:
                fh.write(p.stdout[-20000:] + "\n==== stderr ====\n" + p.stderr[-40000:])
        except OSError:
            pass
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert p.stdout.count("-ok") == world, p.stdout
    return [l for l in p.stdout.splitlines() if l.startswith("CRC ")][0]


def test_sharded_channels_identical_bytes_at_1_and_n_gpus(engine, tmp_path):
    ndev = engine.load().b2d_device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    one = _run(1, tmp_path)
    for world in sorted({2, min(ndev, 8)}):
        assert _run(world, tmp_path) == one, world
