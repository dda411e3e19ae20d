# Phase trace of the ROUND-1 swapped kernel with the fine-grained epilogue stamps added (the reference point for the
# round-2 changes): the old sources are swapped in on the box, built there, traced, and the shipped library is restored.
set -x
cd $GRAFT_REPO_ROOT
cp cadm_b200/libcadm_b200.so /tmp/lib_keep.so; cp cadm_b200/.libcadm_b200.stamp /tmp/stamp_keep
cp cadm_b200/csrc/rollout_tcs.cu /tmp/tcs_keep.cu; cp cadm_b200/csrc/ptx.cuh /tmp/ptx_keep.cuh
cp tools/probes/rollout_tcs_r1_stamped.cu.txt cadm_b200/csrc/rollout_tcs.cu; cp tools/probes/ptx_r1.cuh.txt cadm_b200/csrc/ptx.cuh
timeout -k 5 400 python -m cadm_b200.build 2>&1 | tail -3
timeout -k 5 300 python tools/tcs_diag.py > gpurun_out/r2e_oldtrace.log 2>&1
cp /tmp/tcs_keep.cu cadm_b200/csrc/rollout_tcs.cu; cp /tmp/ptx_keep.cuh cadm_b200/csrc/ptx.cuh
cp /tmp/lib_keep.so cadm_b200/libcadm_b200.so; cp /tmp/stamp_keep cadm_b200/.libcadm_b200.stamp
