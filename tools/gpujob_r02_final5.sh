# last check of the round: the whole GPU suite + smoke + the default bench line on the final tree
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_final.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_smoke.txt
timeout 600 python bench.py > gpurun_out/r02_bench_fir256.json 2> gpurun_out/r02_bench_fir256.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_fir256.json')); e = d.get('e2e') or {}; s = d['secondary']['cic_dec']
print(f"fir256 {d['config']['kernel_path']} value {d['value']:.1f} frac {d['roofline']['frac']:.4f} e2e {e.get('value', 0):.1f} parity {d['parity']['ok']} | cic_dec {s['value']:.1f} frac {s['roofline']['frac']:.4f} parity {s['parity']['ok']} | cpu {d['cpu_baseline']['value']:.2f} on {d['cpu_baseline']['cores']} cores")
PY
