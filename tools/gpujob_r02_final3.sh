# r02 closing evidence on the tree with the overlap-save FIR path (1 GPU): the whole GPU suite, more fuzz seeds, every bench
# line, launch list + ncu --set full captures of fir_ovs (IQ pair and real channels), compute-sanitizer over the new kernel.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu_final.txt
for seed in 7 8 9; do B2D_FUZZ_SEED=$seed timeout 600 python -m pytest tests/test_fir_ovs.py -m gpu -q -k random_q15_family 2>&1 | tail -1; done | tee gpurun_out/r02_ovs_family_fuzz.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_smoke.txt
timeout 600 python bench.py > gpurun_out/r02_bench_fir256.json 2> gpurun_out/r02_bench_fir256.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 300 python bench.py --workload fir1024 --no-cpu --steps 20 --warmup 5 > gpurun_out/r02_bench_fir1024.json 2> gpurun_out/r02_bench_fir1024.err
B2D_FIR_OVS=0 timeout 300 python bench.py --workload fir256 --no-cpu --no-secondary --steps 20 --warmup 5 > gpurun_out/r02_bench_fir256_q15.json 2> /dev/null
B2D_FIR_OVS=0 timeout 300 python bench.py --workload fir1024 --no-cpu --no-e2e --steps 20 --warmup 5 > gpurun_out/r02_bench_fir1024_q15.json 2> /dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_bench_fir*.json')) + ['gpurun_out/r02_bench_reference.json']:
    try: d = json.load(open(f))
    except Exception as e: print(f, 'unreadable'); continue
    if d.get('impl') == 'reference': print(f, 'reference', d['value'], d['cpu_baseline']['cores']); continue
    e = d.get('e2e') or {}; p = d.get('e2e_packed') or {}
    print(f"{f[21:-5]:14s} {d['config']['kernel_path']:8s} value {d['value']:11.1f} frac {d['roofline']['frac']:.4f} e2e {e.get('value', 0):9.1f} ({e.get('frac', 0) or 0:.3f}) packed {p.get('value', 0):9.1f} parity {(d.get('parity') or {}).get('ok')} fp64 {(d['roofline'].get('fp64_pipe') or {}).get('frac')}")
    if 'secondary' in d:
        s = d['secondary']['cic_dec']; e = s.get('e2e') or {}
        print(f"{'  cic_dec':14s} value {s['value']:11.1f} frac {s['roofline']['frac']:.4f} e2e {e.get('value', 0):9.1f} parity {(s.get('parity') or {}).get('ok')}")
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > /dev/null 2>&1
grep -c . gpurun_out/r02_launches_default.csv; grep "fir_ovs\|cic_dec_fast" gpurun_out/r02_launches_default.csv | tail -4
cap() { name=$1; regex=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_${name}_full \
    python bench.py "$@" --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity --no-secondary > gpurun_out/r02_ncu_$name.log 2>&1
}
cap fir_ovs fir_ovs_kernel --workload fir256 --log2n 26
cap fir_ovs_real fir_ovs_kernel --workload fir1024 --log2n 23
ls -la gpurun_out/r02_fir_ovs*_full.ncu-rep
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_fir_ovs.py -x -q -m gpu -k "selection or device_buffers or (every_architecture and 257) or (formats and fmts0)" > gpurun_out/r02_sanitize_ovs_memcheck.log 2>&1; echo "fir_ovs memcheck rc=$?" | tee gpurun_out/r02_sanitize_ovs_summary.txt
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_fir_ovs.py -x -q -m gpu -k "device_buffers or (every_architecture and 257 and SHIFT) or (formats and fmts0 and planar-3)" > gpurun_out/r02_sanitize_ovs_racecheck.log 2>&1; echo "fir_ovs racecheck rc=$?" | tee -a gpurun_out/r02_sanitize_ovs_summary.txt
tail -n 3 gpurun_out/r02_sanitize_ovs_memcheck.log; tail -n 3 gpurun_out/r02_sanitize_ovs_racecheck.log
