# r02 final 8-GPU job: multi-GPU parity test (keeps the ranks' output on failure), default bench at 8 / 4 / 2 GPUs on the final tree
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_multi_gpu_n${N}_final.txt
for n in $N 4 2; do
  [ $n -gt $N ] && continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 \
    > gpurun_out/r02_bench_final_n$n.json 2> gpurun_out/r02_bench_final_n$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_bench_final_n*.json')):
    try: d = json.load(open(f))
    except Exception as e: print(f, 'unreadable', e); continue
    def show(n, m):
        e, p = m.get('e2e') or {}, m.get('e2e_packed') or {}
        print(f"{f[-12:-5]} {n:8s} value {m['value']:11.0f} frac {m['roofline']['frac']:.3f} e2e {e.get('value', 0):8.0f} ({e.get('frac') or 0:.3f}) packed {p.get('value', 0):8.0f} parity {(m.get('parity') or {}).get('mismatches')}/{(m.get('parity') or {}).get('outputs_checked')}")
    show('fir256', d); show('cic_dec', d['secondary']['cic_dec'])
PY
ls gpurun_out | grep failure
