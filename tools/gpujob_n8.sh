# 8-GPU evidence for BASELINE configs[1], [3], [4]: one rank per GPU, channels sharded, NCCL coefficient broadcast
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
python -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/r01_pytest_multi_gpu$N.log 2>&1; tail -2 gpurun_out/r01_pytest_multi_gpu$N.log
for wl in fir256 fir1024 cicfir cic_dec; do
  $TR bench.py --gpus $N --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err
  cut -c1-200 gpurun_out/bench_${wl}_n$N.json
done
