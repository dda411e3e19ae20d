# A/B of the mixed-precision FHADD split (csrc/ptx.cuh, -DCADM_SPLIT_FHADD=1) on the GPU box:
# default build first (bench line), then the variant is built ON the box (same image, nvcc present, ~40 s), must pass the
# tensor-core parity tests and the virtual-rank bit-identity tests, and is benched; the default library is restored at the end.
#   gpurun --timeout 1500 -- 'bash tools/ab_fhadd.sh'
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 5 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err; cat gpurun_out/ab_default.json
cp cadm_b200/libcadm_b200.so /tmp/libcadm_default.so; cp cadm_b200/.libcadm_b200.stamp /tmp/stamp_default 2>/dev/null
CADM_EXTRA_NVCC_FLAGS="-DCADM_SPLIT_FHADD=1" timeout -k 5 300 python -m cadm_b200.build 2>&1 | tail -3
export CADM_EXTRA_NVCC_FLAGS="-DCADM_SPLIT_FHADD=1"          # keeps the stamp valid for the runs below
timeout -k 5 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/ab_fhadd_tests.log; cat gpurun_out/ab_fhadd_tests.log
timeout -k 5 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_fhadd.json 2> gpurun_out/ab_fhadd.err; cat gpurun_out/ab_fhadd.json
unset CADM_EXTRA_NVCC_FLAGS
cp /tmp/libcadm_default.so cadm_b200/libcadm_b200.so; cp /tmp/stamp_default cadm_b200/.libcadm_b200.stamp 2>/dev/null
python - <<'PY'
import json
a, b = (json.load(open(f"gpurun_out/ab_{k}.json")) for k in ("default", "fhadd"))
print("rollout launch ms: default %.4f  fhadd %.4f  (%.1f %%)" % (a["roofline"]["launch_ms"], b["roofline"]["launch_ms"],
      100 * (b["roofline"]["launch_ms"] / a["roofline"]["launch_ms"] - 1)))
PY
