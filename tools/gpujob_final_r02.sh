# r02 evidence job (1 GPU): the whole GPU suite, every bench line, launch list + ncu --set full captures of the dominant
# kernels, compute-sanitizer over the new code, facade throughput.  Results land in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu_final.txt
timeout 600 python bench.py > gpurun_out/r02_bench_fir256.json 2> gpurun_out/r02_bench_fir256.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
for wl in fir1024 cic_intr cicfir polydec polyintr intgdump fir63; do
  timeout 300 python bench.py --workload $wl --no-cpu --steps 20 --warmup 5 > gpurun_out/r02_bench_$wl.json 2> gpurun_out/r02_bench_$wl.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*.json')):
    try: d = json.load(open(f))
    except Exception as e: print(f, 'unreadable'); continue
    if d.get('impl') == 'reference': print(f, 'reference', d['value'], d['cpu_baseline']['cores']); continue
    e = d.get('e2e') or {}; p = d.get('e2e_packed') or {}
    print(f"{f[21:-5]:10s} value {d['value']:11.1f} frac {d['roofline']['frac']:.4f} e2e {e.get('value', 0):9.1f} ({e.get('frac', 0) or 0:.3f}) packed {p.get('value', 0):9.1f} parity {(d.get('parity') or {}).get('ok')}")
    if 'secondary' in d:
        s = d['secondary']['cic_dec']; e = s.get('e2e') or {}
        print(f"{'  cic_dec':10s} value {s['value']:11.1f} frac {s['roofline']['frac']:.4f} e2e {e.get('value', 0):9.1f} ({e.get('frac', 0) or 0:.3f}) parity {(s.get('parity') or {}).get('ok')} cpu {s.get('cpu_baseline', {}).get('value')}")
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > /dev/null 2>&1
tail -3 gpurun_out/r02_launches_default.csv
cap() { name=$1; regex=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_${name}_full \
    python bench.py "$@" --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity --no-secondary > gpurun_out/r02_ncu_$name.log 2>&1
}
cap fir_q15 fir_q15_kernel --workload fir256 --log2n 26
cap upfir_lane upfir_lane_kernel --workload cicfir
cap fir_q24 fir_q24_kernel --workload fir63
cap cic_dec_fast cic_dec_fast_kernel --workload cic_dec --log2n 28
cap polydec_q15 polydec_q15_kernel --workload polydec --log2n 28
ls -la gpurun_out/r02_*_full.ncu-rep
bash tools/facade_throughput.sh run 1048576 4096 > gpurun_out/r02_facade_throughput.jsonl 2>&1
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_zz_reference_quirks.py -x -q -m gpu -k "q24 and in20 or packed_wire or transposed_partial and 16 or mv_avg and mv0 or per_sample" > gpurun_out/r02_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" | tee gpurun_out/r02_sanitize_summary.txt
timeout 600 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "q24 and in20" > gpurun_out/r02_sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02_sanitize_summary.txt
tail -3 gpurun_out/r02_sanitize_memcheck.log gpurun_out/r02_sanitize_racecheck.log
