# r02 job K: inner-loop unroll A/B on the DP2A kernels (variant libraries via B2D_LIBRARY)
mkdir -p gpurun_out
V=ac_dsp_b200/lib/variants
run() { name=$1; wl=$2; lib=$3
  env ${lib:+B2D_LIBRARY=$lib} timeout 200 python bench.py --workload $wl --no-cpu --no-e2e --steps 20 --warmup 5 > gpurun_out/r02_k_$name.json 2> gpurun_out/r02_k_$name.err
  python - gpurun_out/r02_k_$name.json $name <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"{sys.argv[2]:22s} {d['value']:10.1f}  roofline {d['roofline']['frac']:.4f}  parity {d['parity']['ok'] if d.get('parity') else None}")
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
run fir256_base fir256 ""
run fir256_u2 fir256 $PWD/$V/libb200dsp_u2.so
run fir256_u4 fir256 $PWD/$V/libb200dsp_u4.so
run fir256_base2 fir256 ""
run fir1024_base fir1024 ""
run fir1024_u2 fir1024 $PWD/$V/libb200dsp_u2.so
run fir1024_u4 fir1024 $PWD/$V/libb200dsp_u4.so
run fir63_base fir63 ""
run fir63_u2 fir63 $PWD/$V/libb200dsp_q24u2.so
run polydec_base polydec ""
run polydec_u2 polydec $PWD/$V/libb200dsp_decu2.so
