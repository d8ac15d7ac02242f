python -m pytest tests -m gpu -x -q > gpurun_out/r01_pytest_gpu24.log 2>&1; tail -3 gpurun_out/r01_pytest_gpu24.log
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --workload polyintr > gpurun_out/bench_polyintr.json 2> gpurun_out/bench_polyintr.err; cut -c1-220 gpurun_out/bench_polyintr.json
python bench.py > gpurun_out/bench_fir256_c.json 2> gpurun_out/bench_fir256_c.err; cut -c1-220 gpurun_out/bench_fir256_c.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref2.json 2> gpurun_out/bench_ref2.err; cut -c1-300 gpurun_out/bench_ref2.json
