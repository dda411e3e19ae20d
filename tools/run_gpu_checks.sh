set -x
cd $GRAFT_REPO_ROOT
for cand in 200 1000 5000; do for part in 20 100; do
timeout -k 5 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --cand $cand --part $part 2>/dev/null | tail -1 >> gpurun_out/r1_sweep_c5.jsonl
done; done
cat gpurun_out/r1_sweep_c5.jsonl
