# r02 job O: smoke() and the default bench with the e2e byte check
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 600 python bench.py --no-cpu > gpurun_out/r02_o_bench.json 2> gpurun_out/r02_o_bench.err; tail -2 gpurun_out/r02_o_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_o_bench.json'))
for n,m in (('fir256',d),('cic_dec',d['secondary']['cic_dec'])):
    for k in ('e2e','e2e_packed'):
        if m.get(k): print(n,k,'%.0f'%m[k]['value'], m[k].get('frac'), 'bytes equal device path:', m[k].get('output_equals_device_path'))
PY
for wl in fir1024 fir63; do timeout 300 python bench.py --workload $wl --no-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
for k in ('e2e','e2e_packed'): print('$wl',k,'%.0f'%d[k]['value'],d[k].get('output_equals_device_path'))"; done
