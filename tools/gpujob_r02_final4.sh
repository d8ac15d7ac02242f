# r02 closing run on the final tree: the whole GPU suite, the overlap-save family fuzz on more seeds, smoke(), the two FIR bench lines
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_final.txt
for seed in 1 2 3 4 5 6; do B2D_FUZZ_SEED=$seed timeout 600 python -m pytest tests/test_fir_ovs.py -m gpu -q -k random_q15_family 2>&1 | tail -1; done | tee gpurun_out/r02_ovs_family_fuzz.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_smoke.txt
timeout 600 python bench.py > gpurun_out/r02_bench_fir256.json 2> gpurun_out/r02_bench_fir256.err
timeout 300 python bench.py --workload fir1024 --no-cpu --steps 20 --warmup 5 > gpurun_out/r02_bench_fir1024.json 2> gpurun_out/r02_bench_fir1024.err
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench_fir256.json', 'gpurun_out/r02_bench_fir1024.json'):
    d = json.load(open(f)); e = d.get('e2e') or {}; p = d.get('e2e_packed') or {}
    print(f"{f[21:-5]:10s} {d['config']['kernel_path']:8s} value {d['value']:11.1f} frac {d['roofline']['frac']:.4f} e2e {e.get('value', 0):9.1f} ({e.get('frac', 0) or 0:.3f}) packed {p.get('value', 0):9.1f} parity {(d.get('parity') or {}).get('ok')} fp64 {(d['roofline'].get('fp64_pipe') or {}).get('frac')} traffic {d['roofline']['traffic']}")
    if 'secondary' in d:
        s = d['secondary']['cic_dec']
        print(f"  cic_dec value {s['value']:11.1f} frac {s['roofline']['frac']:.4f} parity {(s.get('parity') or {}).get('ok')}")
PY
