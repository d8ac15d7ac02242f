# first GPU contact of the overlap-save FIR path: parity, A/B against the DP2A kernel, one full ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fir_ovs.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_ovs1_pytest.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fir_q15 or per_channel or device_path_and_state or path_is_taken" 2>&1 | tail -8 >> gpurun_out/r02_ovs1_pytest.txt
for ov in 0 1; do for wl in fir256 fir1024; do
  B2D_FIR_OVS=$ov timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --no-secondary --steps 10 --warmup 3 2>gpurun_out/r02_ovs1_$wl_$ov.err | python -c "
import json,sys; d=json.load(sys.stdin); print('$wl ovs=$ov', round(d['value'],1), d['config']['kernel_path'], d['parity'], round(d['ms_per_step'],3))"
done; done > gpurun_out/r02_ovs1_ab.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fir_ovs_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_fir_ovs_full \
  python bench.py --workload fir256 --log2n 26 --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity --no-secondary > gpurun_out/r02_ncu_fir_ovs.log 2>&1
cat gpurun_out/r02_ovs1_pytest.txt gpurun_out/r02_ovs1_ab.txt
