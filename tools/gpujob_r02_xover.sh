mkdir -p gpurun_out
timeout 600 python tools/ovs_crossover.py > gpurun_out/r02_ovs_crossover2.jsonl 2> gpurun_out/r02_ovs_crossover.err
timeout 600 python -m pytest tests/test_fir_ovs.py -m gpu -q -k "selection" 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
cut -c1-220 gpurun_out/r02_ovs_crossover2.jsonl
