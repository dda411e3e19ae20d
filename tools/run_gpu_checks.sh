set -x
cd $GRAFT_REPO_ROOT
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -x -q -k "tiles or tc_gemm" 2>&1 | tail -4
CADM_TC_VARIANT=1 timeout -k 5 300 python tools/tc_trace.py tc3x 2>&1 | tail -4
timeout -k 5 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config C4
timeout -k 5 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config C2 --m 10
