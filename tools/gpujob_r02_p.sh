mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_engine_fuzz.py -m gpu -q -k "q24 or fir_random or cascade" 2>&1 | tail -2
for i in 1 2; do timeout 200 python bench.py --workload fir63 --no-cpu --no-e2e --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('fir63', d['value'], d['roofline']['frac'], d['parity']['ok'])"; done
timeout 200 python bench.py --workload cicfir --no-cpu --no-e2e --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('cicfir', d['value'], d['roofline']['frac'])"
B2D_CICFIR_TWO_STAGE=1 timeout 200 python bench.py --workload cicfir --no-cpu --no-e2e --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('cicfir two-stage', d['value'], d['roofline']['frac'])"
