# Bisect of the round-2 kernel changes on the GPU box: each variant of rollout_tcs.cu under tools/probes/variants/ is swapped
# in, built on the box and timed (tools/tcs_one.py); the shipped sources and library are restored at the end.
set -x
cd $GRAFT_REPO_ROOT
cp cadm_b200/libcadm_b200.so /tmp/lib_keep.so; cp cadm_b200/.libcadm_b200.stamp /tmp/stamp_keep; cp cadm_b200/csrc/rollout_tcs.cu /tmp/tcs_keep.cu
: > gpurun_out/r2h_bisect.log
python tools/tcs_one.py current >> gpurun_out/r2h_bisect.log 2>&1
for v in "$@"; do
  cp tools/probes/variants/$v.cu.txt cadm_b200/csrc/rollout_tcs.cu
  timeout -k 5 400 python -m cadm_b200.build 2>&1 | tail -2
  python tools/tcs_one.py $v >> gpurun_out/r2h_bisect.log 2>&1
done
cp /tmp/tcs_keep.cu cadm_b200/csrc/rollout_tcs.cu; cp /tmp/lib_keep.so cadm_b200/libcadm_b200.so; cp /tmp/stamp_keep cadm_b200/.libcadm_b200.stamp
grep -E "^==|us per rollout|epi :" gpurun_out/r2h_bisect.log
