# r02 closing job (1 GPU): the whole suite, smoke(), the default bench and the reference arm on the final tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_final.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02_bench_fir256.json 2> gpurun_out/r02_bench_fir256.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> /dev/null
for wl in cicfir polyintr fir63 fir1024; do timeout 300 python bench.py --workload $wl --no-cpu --steps 20 --warmup 5 > gpurun_out/r02_bench_$wl.json 2> /dev/null; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*.json')):
    try: d = json.load(open(f))
    except Exception as e: continue
    if d.get('impl') == 'reference': print('reference', d['value'], d['cpu_baseline']['cores']); continue
    e = d.get('e2e') or {}; p = d.get('e2e_packed') or {}
    print(f"{f[21:-5]:10s} value {d['value']:11.1f} frac {d['roofline']['frac']:.4f} e2e {e.get('value', 0):9.1f} ({e.get('frac', 0) or 0:.3f}, bytes ok {e.get('output_equals_device_path')}) packed {p.get('value', 0):9.1f} parity {(d.get('parity') or {}).get('ok')}")
    if 'secondary' in d:
        s = d['secondary']['cic_dec']; e = s.get('e2e') or {}
        print(f"{'  cic_dec':10s} value {s['value']:11.1f} frac {s['roofline']['frac']:.4f} e2e {e.get('value', 0):9.1f} ({e.get('frac', 0) or 0:.3f}) parity {(s.get('parity') or {}).get('ok')} cpu {s.get('cpu_baseline', {}).get('value')}")
PY
