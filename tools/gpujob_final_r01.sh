python -m pytest tests -m gpu -x -q > gpurun_out/r01_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r01_pytest_gpu_final.log
python bench.py > gpurun_out/bench_fir256_final.json 2> gpurun_out/bench_fir256_final.err; cut -c1-200 gpurun_out/bench_fir256_final.json
python bench.py --workload polydec --no-cpu > gpurun_out/bench_polydec_final.json 2>/dev/null; cut -c1-200 gpurun_out/bench_polydec_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fir256_final.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
tail -4 gpurun_out/launches_fir256_final.csv
