# Round 2, first GPU call: the parity tests the round-1 throughput numbers stand on (multi-tile per CTA, C4 decision, C5
# corners), one kept bench line per BASELINE config on the round-1 kernels, the clock64 phase trace, the FHADD A/B and a
# compute-sanitizer pass over the smoke decision.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/r2_call1.sh'
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/r2_gpu.txt
timeout -k 5 600 python -m pytest tests/test_gpu_multitile.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_pytest_multitile.log; cat gpurun_out/r2_pytest_multitile.log
for spec in "C2 1" "C1 1" "C3 1" "C4 1" "C2 10"; do
  set -- $spec
  timeout -k 5 300 python bench.py --config $1 --m $2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_$1_m$2.json 2> gpurun_out/r2a_bench_$1_m$2.err
  cat gpurun_out/r2a_bench_$1_m$2.json
done
timeout -k 5 300 python tools/tcs_sweep.py C2 1 > gpurun_out/r2a_sweep.log 2>&1; tail -14 gpurun_out/r2a_sweep.log
# sanitizer: racecheck (shared-memory hazards) and synccheck (barrier misuse) over one small decision per rollout kernel
for v in 2 1; do
  CADM_TC_VARIANT=$v timeout -k 5 420 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_racecheck_variant$v.log 2>&1; tail -6 gpurun_out/r2a_racecheck_variant$v.log
  CADM_TC_VARIANT=$v timeout -k 5 420 compute-sanitizer --tool synccheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_synccheck_variant$v.log 2>&1; tail -6 gpurun_out/r2a_synccheck_variant$v.log
done
# FHADD A/B: variant built on the box, must pass the tensor-core parity tests, then benched; default library restored
cp cadm_b200/libcadm_b200.so /tmp/libcadm_default.so; cp cadm_b200/.libcadm_b200.stamp /tmp/stamp_default 2>/dev/null
CADM_EXTRA_NVCC_FLAGS="-DCADM_SPLIT_FHADD=1" timeout -k 5 400 python -m cadm_b200.build 2>&1 | tail -3
export CADM_EXTRA_NVCC_FLAGS="-DCADM_SPLIT_FHADD=1"
timeout -k 5 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q -m gpu -k "swapped or tc" 2>&1 | tail -5 > gpurun_out/r2a_fhadd_tests.log; cat gpurun_out/r2a_fhadd_tests.log
timeout -k 5 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_fhadd.json 2> gpurun_out/r2a_bench_fhadd.err; cat gpurun_out/r2a_bench_fhadd.json
unset CADM_EXTRA_NVCC_FLAGS
cp /tmp/libcadm_default.so cadm_b200/libcadm_b200.so; cp /tmp/stamp_default cadm_b200/.libcadm_b200.stamp 2>/dev/null
ls -la gpurun_out
