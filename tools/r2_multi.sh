# Multi-GPU checks (run with gpurun --gpus N): bit-identity of the sharded decision and of get_action() at world = N, the
# peer-timeout behaviour (N = 2), and the bench line with all legs.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_multi.sh 2 <tag>'
set -x
cd $GRAFT_REPO_ROOT
N=${1:-2}
TAG=${2:-r2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout -k 5 400 $RUN tools/multi_gpu_check.py > gpurun_out/${TAG}_multi_gpu_check_${N}gpu.log 2>&1; grep -E "\[|MULTI_GPU_CHECK|Error|error" gpurun_out/${TAG}_multi_gpu_check_${N}gpu.log | tail -20
if [ "$N" = "2" ]; then
  timeout -k 5 200 $RUN tools/peer_timeout_check.py > gpurun_out/${TAG}_peer_timeout_check.log 2>&1; grep -E "raised|PEER_TIMEOUT_CHECK|recover" gpurun_out/${TAG}_peer_timeout_check.log | tail -5
fi
timeout -k 5 400 $RUN bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; cat gpurun_out/${TAG}_bench_${N}gpu.json; tail -3 gpurun_out/${TAG}_bench_${N}gpu.err
timeout -k 5 300 $RUN bench.py --gpus $N --steps 3 --warmup 1 --impl reference > gpurun_out/${TAG}_bench_reference_${N}gpu.json 2>/dev/null; cat gpurun_out/${TAG}_bench_reference_${N}gpu.json
