# r02 job G: fir_q24 with aligned staging, upfir launch-bounds A/B, e2e chunk-size sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_engine_fuzz.py -m gpu -q -k "q24 or q15_path or cascade or fir_random or cicfir" 2>&1 | tail -5
run() { name=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --no-cpu --no-e2e --no-parity --steps 20 --warmup 5 > gpurun_out/r02_g_$name.json 2> gpurun_out/r02_g_$name.err
  python - gpurun_out/r02_g_$name.json $name <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"{sys.argv[2]:28s} {d['value']:10.1f} {d['unit']}  roofline {d['roofline']['frac']:.3f}  path {d['config']['kernel_path']}")
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
run fir63_q24 fir63 B2D_X=0
run cicfir_default cicfir B2D_X=0
run cicfir_lb6 cicfir B2D_UPFIR_LB6=1
run cicfir_two_stage cicfir B2D_CICFIR_TWO_STAGE=1
for mb in 8 16 24 48 96; do
  B2D_PIPE_CHUNK_BYTES=$((mb*1048576)) timeout 200 python bench.py --no-cpu --no-parity --no-secondary --steps 5 --warmup 3 > gpurun_out/r02_g_e2e_chunk$mb.json 2> gpurun_out/r02_g_e2e_chunk$mb.err
  python - gpurun_out/r02_g_e2e_chunk$mb.json $mb <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(f"chunk {sys.argv[2]:>3s} MiB  e2e {d['e2e']['value']:8.1f} ({d['e2e'].get('frac',0):.3f})  packed {d['e2e_packed']['value']:8.1f} ({d['e2e_packed'].get('frac',0):.3f})")
PY
done
