# r02 job F: q24 parity, upfir epilogue A/B on one box, ncu captures of fir_q24 and upfir_lane
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "q24 or q15_path" 2>&1 | tail -5
run() { name=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --no-cpu --no-e2e --no-parity --steps 20 --warmup 5 > gpurun_out/r02_f_$name.json 2> gpurun_out/r02_f_$name.err
  python - gpurun_out/r02_f_$name.json $name <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"{sys.argv[2]:28s} {d['value']:10.1f} {d['unit']}  roofline {d['roofline']['frac']:.3f}  path {d['config']['kernel_path']}")
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
run cicfir_epi1 cicfir B2D_UPFIR_EPI=1
run cicfir_epi0 cicfir B2D_UPFIR_EPI=0
run cicfir_epi1_nopeel cicfir B2D_UPFIR_EPI=1 B2D_UPFIR_PEEL=0
run cicfir_epi0_nopeel cicfir B2D_UPFIR_EPI=0 B2D_UPFIR_PEEL=0
run cicfir_epi1_again cicfir B2D_UPFIR_EPI=1
run polyintr_epi1 polyintr B2D_UPFIR_EPI=1
run polyintr_epi0 polyintr B2D_UPFIR_EPI=0
timeout 200 ncu --set full --clock-control none --import-source on -k regex:fir_q24_kernel --launch-skip 1 --launch-count 1 -f \
  -o gpurun_out/r02_fir_q24_full python bench.py --workload fir63 --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/r02_ncu_fir_q24.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:upfir_lane_kernel --launch-skip 1 --launch-count 1 -f \
  -o gpurun_out/r02_upfir_lane_full python bench.py --workload cicfir --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/r02_ncu_upfir.log 2>&1
ls -la gpurun_out/*.ncu-rep
