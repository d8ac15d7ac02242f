mkdir -p gpurun_out
V=ac_dsp_b200/lib/variants
run() { name=$1; wl=$2; lib=$3
  env ${lib:+B2D_LIBRARY=$lib} timeout 200 python bench.py --workload $wl --no-cpu --no-e2e --no-parity --steps 20 --warmup 5 > gpurun_out/r02_m_$name.json 2> gpurun_out/r02_m_$name.err
  python - gpurun_out/r02_m_$name.json $name <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"{sys.argv[2]:22s} {d['value']:10.1f}  roofline {d['roofline']['frac']:.4f}")
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
run cicfir_base cicfir ""
run cicfir_up1 cicfir $PWD/$V/libb200dsp_up1.so
run cicfir_up4 cicfir $PWD/$V/libb200dsp_up4.so
run cicfir_up8 cicfir $PWD/$V/libb200dsp_up8.so
run polyintr_base polyintr ""
run polyintr_up4 polyintr $PWD/$V/libb200dsp_up4.so
run polyintr_up8 polyintr $PWD/$V/libb200dsp_up8.so
run fir256_base fir256 ""
run fir256_p8 fir256 $PWD/$V/libb200dsp_q15p8.so
run fir256_p2 fir256 $PWD/$V/libb200dsp_q15p2.so
