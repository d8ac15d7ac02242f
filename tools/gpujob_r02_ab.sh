# Round-2 opening GPU job: A/B the switches prepared at the end of round 1 (profiles/r01_source_level_notes.md).
#   gpurun --timeout 600 -- 'bash tools/gpujob_r02_ab.sh'
# Each variant first has to pass the parity tests of the kernels it touches, then is timed with bench.py (device-resident
# inputs, no CPU arm); lines land in gpurun_out/r02_ab_*.json.
mkdir -p gpurun_out
K="cascade or cicfir or poly_intr or polyintr"
run() {   # name, env assignments...
  name=$1; shift
  echo "== $name: $*" | tee -a gpurun_out/r02_ab_summary.txt
  env "$@" timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" 2>&1 | tail -1 | tee -a gpurun_out/r02_ab_summary.txt
  for wl in cicfir polyintr; do
    env "$@" timeout 120 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/r02_ab_${name}_${wl}.json 2> gpurun_out/r02_ab_${name}_${wl}.err
    python - gpurun_out/r02_ab_${name}_${wl}.json <<'PY' | tee -a gpurun_out/r02_ab_summary.txt
import json, sys
d = json.load(open(sys.argv[1]))
print(f"   {d['config']['workload'][:40]:40s} {d['value']:10.1f} {d['unit']}  roofline {d['roofline']['frac']:.3f}")
PY
  done
}
# the random-format engine sweep written at the end of round 1 (never run on a GPU yet)
for seed in 0 1 2 3 4 5; do
  echo "== engine fuzz seed $seed" | tee -a gpurun_out/r02_ab_summary.txt
  B2D_FUZZ_SEED=$seed B2D_ENGINE_FUZZ=1 timeout 900 python -m pytest tests/test_zz_engine_fuzz.py tests/test_zz_reference_quirks.py -m gpu -q -rf 2>&1 | tail -40 | tee gpurun_out/r02_engine_fuzz_seed$seed.txt | tail -3 | tee -a gpurun_out/r02_ab_summary.txt
done
run base     B2D_UPFIR_WAVES=1
run waves2   B2D_UPFIR_WAVES=2
run waves4   B2D_UPFIR_WAVES=4
run waves8   B2D_UPFIR_WAVES=8
run waves16  B2D_UPFIR_WAVES=16
run peel     B2D_UPFIR_PEEL=1
run peel_w4  B2D_UPFIR_PEEL=1 B2D_UPFIR_WAVES=4
run peel_w8  B2D_UPFIR_PEEL=1 B2D_UPFIR_WAVES=8
# source-level capture of the headline kernel (not taken in round 1): where do the 4 % non-IDP slots of the DP2A pipe go?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:fir_q15_kernel --launch-skip 1 --launch-count 1 -f \
  -o gpurun_out/r02_fir_q15_full python bench.py --workload fir256 --log2n 26 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02_ncu_fir_q15.log 2>&1
# e2e at N GPUs is host-bound (round 1: 3.28 G IQ samples/s at 1 GPU, 3.94 G at 2, 5.30 G at 8).  A/B on a multi-GPU box:
#   gpurun --gpus 8 --timeout 900 -- 'for v in 0 1; do B2D_HOST_NUMA=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
#     --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_e2e_n8_numa$v.json; done'
