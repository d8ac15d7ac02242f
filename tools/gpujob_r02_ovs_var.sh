# A/B of the fir_ovs kernel variants (B2D_OVS_VARIANT) on both FIR workloads, device-resident
mkdir -p gpurun_out
for v in 0 2 4 6 8 10 12 14; do for wl in fir256 fir1024; do
  B2D_OVS_VARIANT=$v timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --no-secondary --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('variant $v $wl', round(d['value'],1), d['config']['kernel_path'], d['parity']['ok'], round(d['ms_per_step'],3))"
done; done > gpurun_out/r02_ovs_variants2.txt 2>&1
cat gpurun_out/r02_ovs_variants2.txt
