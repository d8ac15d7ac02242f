# r02 8-GPU job: host-link ceilings at 1/2/4/8 GPUs, multi-GPU parity, the default bench at N=8 with and without NUMA-local host buffers.
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
timeout 300 tools/bin/ubench_pcie 256 8 > gpurun_out/r02_ubench_pcie_n$N.jsonl 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_multi_gpu_n$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_default_n$N.json 2> gpurun_out/r02_bench_default_n$N.err
B2D_HOST_NUMA=1 timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-secondary --no-parity > gpurun_out/r02_bench_numa1_n$N.json 2> gpurun_out/r02_bench_numa1_n$N.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*_n8.json')):
    try: d=json.load(open(f))
    except Exception as e: print(f, 'unreadable', e); continue
    def show(n,m):
        print(f, n, 'value %.0f' % m['value'], 'frac %.3f' % m['roofline']['frac'], 'parity', m.get('parity'))
        for k in ('e2e','e2e_packed'):
            if m.get(k): print('   ',k, '%.0f' % m[k]['value'], m[k].get('frac'), m[k].get('host_numa_alloc'))
    show('fir256', d)
    if 'secondary' in d: show('cic_dec', d['secondary']['cic_dec'])
PY
grep -h '"pattern": "\(d2h\|h2d\|fir256\|fir256p\)", "h2d_source": "default"' gpurun_out/r02_ubench_pcie_n$N.jsonl | cut -c1-200
