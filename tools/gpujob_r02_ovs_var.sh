# A/B of fir_ovs build variants (tools/build_variant.sh) on fir256, device-resident
mkdir -p gpurun_out
for wl in fir256 fir1024; do for v in base pf2 stcs hc8; do
  lib=""; [ $v != base ] && lib=$PWD/ac_dsp_b200/lib/variants/libb200dsp_$v.so
  B2D_LIBRARY=$lib timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --no-secondary --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$v $wl', round(d['value'],1), d['config']['kernel_path'], d['parity']['ok'], round(d['ms_per_step'],3))"
done; done > gpurun_out/r02_ovs_var5.txt 2>&1
cat gpurun_out/r02_ovs_var5.txt
