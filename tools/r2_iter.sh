# Kernel iteration check on the GPU box: parity of the swapped-operand kernel (oracle comparisons + bit-identity across
# shardings), the multi-tile tests, the clock64 phase trace of CTA 0 and one bench line.  Everything lands in gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/r2_iter.sh <tag>'
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-iter}
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_envs.py tests/test_gpu_widening.py -x -q -m gpu -k "not tiles and not fp32" 2>&1 | tail -8 > gpurun_out/${TAG}_parity.log; cat gpurun_out/${TAG}_parity.log
timeout -k 5 300 python -m pytest tests/test_gpu_multitile.py -x -q -m gpu -k "rollout or c4" 2>&1 | tail -8 > gpurun_out/${TAG}_multitile.log; cat gpurun_out/${TAG}_multitile.log
timeout -k 5 300 python tools/tcs_sweep.py C2 1 > gpurun_out/${TAG}_sweep.log 2>&1; tail -16 gpurun_out/${TAG}_sweep.log
timeout -k 5 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench_C2.err; cat gpurun_out/${TAG}_bench_C2.json
