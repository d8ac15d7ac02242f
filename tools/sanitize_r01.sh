# compute-sanitizer over the kernel tests: memcheck on every kernel family, racecheck on the
# shared-memory staged ones.  Slow (10-50x): subsets only.  Logs -> gpurun_out/, summaries copied to profiles/.
set -x
CS=/usr/local/cuda/bin/compute-sanitizer
K='smoke'
timeout 900 $CS --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/sanitize_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?" | tee -a gpurun_out/sanitize_summary.txt
timeout 1500 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "intg_dump_vector or cascade_channels or poly_intr_random or q15_iq or cic_multichannel or poly_dec_ddc or reg_share_random" > gpurun_out/sanitize_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?" | tee -a gpurun_out/sanitize_summary.txt
timeout 900 $CS --tool racecheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/sanitize_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?" | tee -a gpurun_out/sanitize_summary.txt
for f in gpurun_out/sanitize_*.log; do tail -n 4 $f; done
