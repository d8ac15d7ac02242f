mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_engine_fuzz.py tests/test_zz_reference_quirks.py -m gpu -q 2>&1 | tail -2
for wl in fir256 fir256 fir1024; do timeout 200 python bench.py --workload $wl --no-cpu --no-e2e --no-secondary --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$wl', d['value'], d['roofline']['frac'], d['roofline']['int_pipe']['frac'], d['parity']['ok'])"; done
timeout 200 ncu --set full --clock-control none -k regex:fir_q15_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_fir_q15_one_full python bench.py --workload fir256 --log2n 26 --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > /dev/null 2>&1
