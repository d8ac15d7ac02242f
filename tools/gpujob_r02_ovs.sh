# overlap-save FIR path: parity, A/B against the DP2A kernel, optional full ncu capture.  usage: gpujob_r02_ovs.sh TAG [ncu] [quick]
TAG=${1:-x}; mkdir -p gpurun_out
if [ "$3" != "quick" ]; then
timeout 900 python -m pytest tests/test_fir_ovs.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_ovs_${TAG}_pytest.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fir_q15 or per_channel or device_path_and_state or path_is_taken" 2>&1 | tail -8 >> gpurun_out/r02_ovs_${TAG}_pytest.txt
fi
for wl in fir256 fir1024; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --no-secondary --steps 10 --warmup 3 2>gpurun_out/r02_ovs_${TAG}_$wl.err | python -c "
import json,sys; d=json.load(sys.stdin); print('$wl', round(d['value'],1), d['config']['kernel_path'], d['parity']['ok'], round(d['ms_per_step'],3), d['clocks'])"
done > gpurun_out/r02_ovs_${TAG}_ab.txt 2>&1
if [ "$2" = "ncu" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fir_ovs_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_fir_ovs_${TAG}_full \
  python bench.py --workload fir256 --log2n 26 --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity --no-secondary > gpurun_out/r02_ncu_fir_ovs.log 2>&1
fi
cat gpurun_out/r02_ovs_${TAG}_pytest.txt gpurun_out/r02_ovs_${TAG}_ab.txt
