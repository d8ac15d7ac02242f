# r02 job C: full GPU suite after the runtime split / wire format / TRANSPOSED, then the default bench (both targets) and the reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu_c.txt
timeout 600 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -3 gpurun_out/r02_bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
def show(n,m):
    print(n, 'value %.0f' % m['value'], 'ms %.3f' % m['ms_per_step'], 'frac %.3f' % m['roofline']['frac'], 'parity', m['parity'])
    for k in ('e2e','e2e_packed'):
        if m.get(k): print('   ',k, '%.0f' % m[k]['value'], m[k].get('frac'), m[k]['d2h_bytes_per_step'], m[k]['h2d_bytes_per_step'])
    print('    cpu', m.get('cpu_baseline',{}).get('value'), m.get('cpu_baseline',{}).get('cores'))
show('fir256', d); show('cic_dec', d['secondary']['cic_dec'])
PY
