# A/B of the fir_ovs raw-carry switch on fir256 (device-resident) + a short parity run under it
mkdir -p gpurun_out
for v in 0 1 0 1; do
  B2D_OVS_RAW=$v timeout 300 python bench.py --workload fir256 --no-cpu --no-e2e --no-secondary --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('raw $v fir256', round(d['value'],1), d['config']['kernel_path'], d['parity']['ok'], round(d['ms_per_step'],3))"
done > gpurun_out/r02_ovs_raw_ab.txt 2>&1
B2D_OVS_RAW=1 timeout 600 python -m pytest tests/test_fir_ovs.py -m gpu -q -k "device_buffers or selection" 2>&1 | tail -2 >> gpurun_out/r02_ovs_raw_ab.txt
cat gpurun_out/r02_ovs_raw_ab.txt
