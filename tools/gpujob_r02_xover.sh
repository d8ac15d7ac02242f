mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ovs_crossover_launches.csv python tools/ovs_crossover_kernels.py > gpurun_out/r02_ovs_crossover_kernels.log 2>&1
grep -c . gpurun_out/r02_ovs_crossover_launches.csv
