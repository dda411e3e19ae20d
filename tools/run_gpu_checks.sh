set -x
cd $GRAFT_REPO_ROOT
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_envs.py -x -q -k "C3 or C4 or golden or session or drop_in" 2>&1 | tail -4
timeout -k 5 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config C3
