"""N > 1 host logic on CPU: two gloo ranks (the data path itself has no collective; NCCL carries only the taps)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, hashlib
sys.path.insert(0, os.environ["B2D_ROOT"])
import numpy as np
import torch.distributed as dist
from ac_dsp_b200 import parallel as P
rank, world, local = P.rank_world()
dist.init_process_group("gloo")
assert dist.get_world_size() == world == 2
# 1. the rendezvous that carries the communicator id: created on rank 0 only, identical bytes everywhere
made = []
def make_id():
    made.append(1)
    return bytes(range(128))
uid = P.exchange_unique_id(make_id, rank)
assert uid == bytes(range(128)) and len(made) == (1 if rank == 0 else 0)
# 2. channel sharding: 64 channels (BASELINE config 4) and a ragged count
for C in (64, 8, 5, 1):
    mine = P.local_channels(C, rank, world)
    got = [None, None]
    dist.all_gather_object(got, mine)
    assert sorted(got[0] + got[1]) == list(range(C)), got
    assert all(P.owner_of(c, world) == (rank, i) for i, c in enumerate(mine))
    assert P.check_partition(C, world)
# 3. every rank filters its own channels; the job result is the concatenation, independent of world size
rng = np.random.default_rng(7)
x = rng.integers(-100, 100, size=(5, 64))
h = rng.integers(-9, 9, size=4) if rank == 0 else None
box = [h]
dist.broadcast_object_list(box, src=0)          # stands in for the ncclBroadcast inside b2d_fir_load
h = box[0]
part = {c: np.convolve(x[c], h)[:64].tolist() for c in P.local_channels(5, rank, world)}
allp = [None, None]
dist.all_gather_object(allp, part)
merged = {**allp[0], **allp[1]}
assert all(merged[c] == np.convolve(x[c], h)[:64].tolist() for c in range(5))
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("rank%d-ok\n" % rank)
sys.stdout.flush()
'''


def _torchrun(args, env=None, timeout=600):
    import socket
    e = dict(os.environ)
    e.update(env or {})
    with socket.socket() as sk:          # a free rendezvous port
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_two_rank_host_logic_gloo(engine, tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    p = _torchrun([str(w)], env={"B2D_ROOT": ROOT})
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert p.stdout.count("-ok") == 2, p.stdout


def test_reference_arm_under_torchrun_prints_one_line(oracle):
    """bench.py --impl reference with N = 2: rank 0 alone runs and prints; the other rank exits 0 without work."""
    p = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--ref-seconds", "0.3"])
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0
