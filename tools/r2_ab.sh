# A/B of a compile-time switch on the GPU box: the shipped library first, then a variant built ON the box with extra nvcc
# flags (same image, ~40 s), each through the swapped-kernel parity tests, the phase trace and one bench line.
#   gpurun --timeout 1200 -- 'bash tools/r2_ab.sh <tag> "<extra nvcc flags>"'
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-ab}
FLAGS="$2"
mkdir -p gpurun_out
run_one() {
  timeout -k 5 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_envs.py tests/test_gpu_widening.py -x -q -m gpu -k "not tiles and not fp32" 2>&1 | tail -6 > gpurun_out/$1_parity.log; cat gpurun_out/$1_parity.log
  timeout -k 5 300 python tools/tcs_sweep.py C2 1 > gpurun_out/$1_sweep.log 2>&1; tail -16 gpurun_out/$1_sweep.log
  timeout -k 5 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/$1_bench_C2.json 2> gpurun_out/$1_bench_C2.err; cat gpurun_out/$1_bench_C2.json
}
run_one ${TAG}_a
if [ -n "$FLAGS" ]; then
  cp cadm_b200/libcadm_b200.so /tmp/libcadm_default.so; cp cadm_b200/.libcadm_b200.stamp /tmp/stamp_default 2>/dev/null
  export CADM_EXTRA_NVCC_FLAGS="$FLAGS"
  timeout -k 5 400 python -m cadm_b200.build 2>&1 | tail -3
  run_one ${TAG}_b
  unset CADM_EXTRA_NVCC_FLAGS
  cp /tmp/libcadm_default.so cadm_b200/libcadm_b200.so; cp /tmp/stamp_default cadm_b200/.libcadm_b200.stamp 2>/dev/null
fi
