# Quick GPU checks while iterating on a kernel (run through gpurun): the tensor-core self-tests, the parity tests of the
# swapped-operand kernel, and the tile-rows sweep with the clock64 phase trace of CTA 0.  Everything lands in gpurun_out/.
set -x
cd $GRAFT_REPO_ROOT
timeout -k 5 300 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -4
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -x -q -k "swapped" 2>&1 | tail -4
timeout -k 5 400 python tools/tcs_sweep.py C2 1 2>&1 | tail -12
