# warp-level barriers between the group-local passes: bench, parity, racecheck
mkdir -p gpurun_out
for wl in fir256 fir1024 fir256 fir1024; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --no-secondary --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('syncwarp $wl', round(d['value'],1), d['config']['kernel_path'], d['parity']['ok'], round(d['ms_per_step'],3))"
done > gpurun_out/r02_ovs_syncwarp.txt 2>&1
timeout 600 python -m pytest tests/test_fir_ovs.py -m gpu -q 2>&1 | tail -2 >> gpurun_out/r02_ovs_syncwarp.txt
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_fir_ovs.py -x -q -m gpu -k "device_buffers or (every_architecture and 257 and SHIFT) or (formats and fmts0 and planar-3)" > gpurun_out/r02_sanitize_ovs_racecheck.log 2>&1; echo "fir_ovs racecheck rc=$?" >> gpurun_out/r02_ovs_syncwarp.txt
tail -n 2 gpurun_out/r02_sanitize_ovs_racecheck.log >> gpurun_out/r02_ovs_syncwarp.txt
cat gpurun_out/r02_ovs_syncwarp.txt
