# BASELINE configs[4] cells on N GPUs (strong: the cell's candidates split over the ranks).  gpurun --gpus N -- 'bash tools/r2_sweep_multi.sh N tag'
set -x
cd $GRAFT_REPO_ROOT
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621"
rm -f gpurun_out/${TAG}_sweep_c5_${N}gpu.jsonl
for cp in "1000 100" "5000 20" "5000 100"; do
  set -- $cp
  timeout -k 5 300 $RUN bench.py --gpus $N --config C2 --cand $1 --part $2 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' >> gpurun_out/${TAG}_sweep_c5_${N}gpu.jsonl
done
wc -l gpurun_out/${TAG}_sweep_c5_${N}gpu.jsonl
