mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:upfir_f64_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_upfir_f64_full \
  python bench.py --workload cicfir --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity --no-secondary > gpurun_out/r02_ncu_upfir_f64.log 2>&1
ls -la gpurun_out/r02_upfir_f64_full.ncu-rep
