# r02 job H: fir_q24 with the lead-shifted staging; host-link rate when the pinned buffer is as large as the e2e call's (2 GiB)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_engine_fuzz.py -m gpu -q -k "q24 or q15_path or cascade or fir_random" 2>&1 | tail -3
run() { name=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --no-cpu --no-e2e --no-parity --steps 20 --warmup 5 > gpurun_out/r02_h_$name.json 2> gpurun_out/r02_h_$name.err
  python - gpurun_out/r02_h_$name.json $name <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"{sys.argv[2]:28s} {d['value']:10.1f} {d['unit']}  roofline {d['roofline']['frac']:.3f}  path {d['config']['kernel_path']}")
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
run fir63_q24 fir63 B2D_X=0
run cicfir_two_stage cicfir B2D_CICFIR_TWO_STAGE=1
timeout 300 tools/bin/ubench_pcie 2048 3 > gpurun_out/r02_ubench_pcie_2g.jsonl 2>&1; grep -h '"h2d_source": "default"' gpurun_out/r02_ubench_pcie_2g.jsonl | cut -c1-170
timeout 200 ncu --set full --clock-control none --import-source on -k regex:fir_q24_kernel --launch-skip 1 --launch-count 1 -f \
  -o gpurun_out/r02_fir_q24_full python bench.py --workload fir63 --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/r02_ncu_fir_q24.log 2>&1
