# r02 job D: full GPU suite (no -x), e2e with 24 MiB chunks
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r02_pytest_gpu_d.txt
timeout 600 python bench.py --no-cpu > gpurun_out/r02_bench_default_d.json 2> gpurun_out/r02_bench_default_d.err; tail -3 gpurun_out/r02_bench_default_d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default_d.json'))
def show(n,m):
    print(n, 'value %.0f' % m['value'], 'ms %.3f' % m['ms_per_step'], 'frac %.3f' % m['roofline']['frac'], 'parity', m['parity']['ok'])
    for k in ('e2e','e2e_packed'):
        if m.get(k): print('   ',k, '%.0f' % m[k]['value'], m[k].get('frac'), m[k]['d2h_bytes_per_step'], m[k]['h2d_bytes_per_step'])
show('fir256', d); show('cic_dec', d['secondary']['cic_dec'])
PY
