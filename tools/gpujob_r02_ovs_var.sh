# A/B of the fused phase-E / epilogue switch, smoke(), the reg_share case through both evaluations
mkdir -p gpurun_out
for v in 0 1 0 1; do for wl in fir256 fir1024; do
  B2D_OVS_FUSE=$v timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --no-secondary --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('fuse $v $wl', round(d['value'],1), d['config']['kernel_path'], d['parity']['ok'], round(d['ms_per_step'],3))"
done; done > gpurun_out/r02_ovs_fuse_ab.txt 2>&1
B2D_OVS_FUSE=1 timeout 600 python -m pytest tests/test_fir_ovs.py -m gpu -q 2>&1 | tail -2 >> gpurun_out/r02_ovs_fuse_ab.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "reg_share_random" 2>&1 | tail -2 >> gpurun_out/r02_ovs_fuse_ab.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> gpurun_out/r02_ovs_fuse_ab.txt
cat gpurun_out/r02_ovs_fuse_ab.txt
