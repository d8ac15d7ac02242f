# r02 job I: full GPU suite on the current tree, facade throughput, fir63 with the templated staging
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu_i.txt
bash tools/facade_throughput.sh run 1048576 4096 > gpurun_out/r02_facade_throughput.jsonl 2>&1; cat gpurun_out/r02_facade_throughput.jsonl
timeout 200 python bench.py --workload fir63 --no-cpu --no-e2e --no-parity --steps 20 --warmup 5 > gpurun_out/r02_i_fir63.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_i_fir63.json')); print('fir63', d['value'], d['roofline']['frac'], d['config']['kernel_path'])"
