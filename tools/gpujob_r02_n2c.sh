# 2-GPU check of the final tree: multi-GPU parity test + the default bench under torchrun
mkdir -p gpurun_out
N=${1:-2}
[ "$2" = "notest" ] || timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_multi_gpu_n${N}_ovs.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/r02_bench_ovs_n$N.json 2> gpurun_out/r02_bench_ovs_n$N.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02_bench_ovs_n$N.json'))
def show(n, m):
    e, p = m.get('e2e') or {}, m.get('e2e_packed') or {}
    print(f"n$N {n:8s} {m.get('config',{}).get('kernel_path')} value {m['value']:11.0f} frac {m['roofline']['frac']:.3f} e2e {e.get('value', 0):8.0f} ({e.get('frac') or 0:.3f}) packed {p.get('value', 0):8.0f} parity {(m.get('parity') or {}).get('mismatches')}/{(m.get('parity') or {}).get('outputs_checked')} clocks {m.get('clocks')}")
show('fir256', d); show('cic_dec', d['secondary']['cic_dec'])
PY
