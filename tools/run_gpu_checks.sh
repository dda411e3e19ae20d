set -x
cd $GRAFT_REPO_ROOT
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_envs.py tests/test_golden.py -x -q -k "swapped or envs or tc3x or golden" 2>&1 | tail -4
timeout -k 5 400 python tools/tcs_sweep.py C2 1 2>&1 | tail -8
