set -x
cd $GRAFT_REPO_ROOT
timeout -k 5 300 python -m pytest tests/test_gpu_tc.py -x -q -k tcs 2>&1 | tail -15 > gpurun_out/r1_tc_tests.log; cat gpurun_out/r1_tc_tests.log
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -x -q -k "swapped" 2>&1 | tail -25 > gpurun_out/r1_parity_swapped.log; cat gpurun_out/r1_parity_swapped.log
timeout -k 5 400 python tools/tcs_sweep.py C2 1 > gpurun_out/r1_sweep.log 2>&1; cat gpurun_out/r1_sweep.log
