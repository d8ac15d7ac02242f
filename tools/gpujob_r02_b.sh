# r02 job B: full GPU parity suite on the tree with TRANSPOSED partial sums + default engine fuzz, pipe and PCIe ubenches.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu_b.txt
for seed in 11 12 13; do
  B2D_FUZZ_SEED=$seed timeout 600 python -m pytest tests/test_zz_engine_fuzz.py -m gpu -q 2>&1 | tail -4 | tee -a gpurun_out/r02_pytest_gpu_b.txt
done
timeout 120 tools/bin/ubench_pipes > gpurun_out/r02_ubench_pipes.jsonl 2>&1
timeout 300 tools/bin/ubench_pcie 256 12 > gpurun_out/r02_ubench_pcie_n1.jsonl 2>&1
tail -8 gpurun_out/r02_ubench_pipes.jsonl; cat gpurun_out/r02_ubench_pcie_n1.jsonl
