mkdir -p gpurun_out
V=ac_dsp_b200/lib/variants
run() { name=$1; wl=$2; lib=$3
  env ${lib:+B2D_LIBRARY=$lib} timeout 200 python bench.py --workload $wl --no-cpu --no-e2e --steps 20 --warmup 5 > gpurun_out/r02_l_$name.json 2> gpurun_out/r02_l_$name.err
  python - gpurun_out/r02_l_$name.json $name <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"{sys.argv[2]:22s} {d['value']:10.1f}  roofline {d['roofline']['frac']:.4f}  parity {d['parity']['ok'] if d.get('parity') else None}")
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
run fir256_u4 fir256 $PWD/$V/libb200dsp_u4.so
run fir256_u8 fir256 $PWD/$V/libb200dsp_u8.so
run fir256_u16 fir256 $PWD/$V/libb200dsp_u16.so
run fir1024_u8 fir1024 $PWD/$V/libb200dsp_u8.so
run fir1024_u16 fir1024 $PWD/$V/libb200dsp_u16.so
