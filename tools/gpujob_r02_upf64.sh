# FP64 form of the fused cascade kernel (taps as kernel parameters): parity + A/B against the DP2A form
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cascade or checkpoint_resume" 2>&1 | tail -3 > gpurun_out/r02_upf64c.txt
for v in 0 1 0 1; do
  B2D_UPFIR_F64=$v timeout 300 python bench.py --workload cicfir --no-cpu --no-e2e --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('f64=$v cicfir', round(d['value'],1), d['config']['kernel_path'], (d.get('parity') or {}).get('ok'), round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
done >> gpurun_out/r02_upf64c.txt 2>&1
cat gpurun_out/r02_upf64c.txt
