# r02 job J: ac_mv_avg on the GPU (fixtures, facade, random formats), plus the whole suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r02_pytest_gpu_j.txt
for seed in 21 22 23; do B2D_FUZZ_SEED=$seed timeout 600 python -m pytest tests/test_zz_engine_fuzz.py -m gpu -q 2>&1 | tail -2; done
