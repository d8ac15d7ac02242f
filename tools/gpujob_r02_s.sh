mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_s_pytest.txt
for seed in 31 32 33 34 35 36; do B2D_FUZZ_SEED=$seed timeout 600 python -m pytest tests/test_zz_engine_fuzz.py -m gpu -q 2>&1 | tail -4; done > gpurun_out/r02_s_fuzz.txt
B2D_GENERIC_I128=1 timeout 900 python -m pytest tests/test_zz_engine_fuzz.py tests/test_gpu_parity.py -m gpu -q -k "random or reference_outputs or mv_avg" 2>&1 | tail -2 > gpurun_out/r02_s_i128.txt
for v in 0 1; do B2D_GENERIC_I128=$v B2D_FORCE_GENERIC=1 timeout 300 python bench.py --workload fir63 --log2n 24 --no-cpu --no-e2e --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('fir63 generic i128=$v', d['value'], d['config']['kernel_path'], d['parity']['ok'])"; done > gpurun_out/r02_s_generic_ab.txt
cat gpurun_out/r02_s_*.txt
