# r02 job E: fir_q24 (new), upfir defaults (waves 8, peel on 3 planes, 32-bit epilogue), full suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r02_pytest_gpu_e.txt
run() { name=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --no-cpu --no-e2e --no-parity --steps 20 --warmup 5 > gpurun_out/r02_e_$name.json 2> gpurun_out/r02_e_$name.err
  python - gpurun_out/r02_e_$name.json $name <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"{sys.argv[2]:28s} {d['value']:10.1f} {d['unit']}  roofline {d['roofline']['frac']:.3f}  path {d['config']['kernel_path']}")
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
run fir63_q24 fir63 B2D_X=0
run fir63_wide fir63 B2D_FORCE_GENERIC=2
run cicfir_default cicfir B2D_X=0
run cicfir_nopeel cicfir B2D_UPFIR_PEEL=0
run cicfir_w16 cicfir B2D_UPFIR_WAVES=16
run cicfir_w4 cicfir B2D_UPFIR_WAVES=4
run cicfir_two_stage cicfir B2D_CICFIR_TWO_STAGE=1
run polyintr_default polyintr B2D_X=0
run polyintr_w16 polyintr B2D_UPFIR_WAVES=16
run polydec polydec B2D_X=0
run fir1024 fir1024 B2D_X=0
run cic_intr cic_intr B2D_X=0
run intgdump intgdump B2D_X=0
