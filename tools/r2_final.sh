# Round-2 final single-GPU record: whole GPU suite, smoke, one bench line per BASELINE config (C1-C4, C2 with m = 10, the C5
# sweep cells), the launch list of the bench command, one `ncu --set full` capture of the rollout kernel of C2 and of C4, the
# training-step timing and a memcheck pass over the hand-written trainer.  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 2400 -- 'bash tools/r2_final.sh r2f'
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-r2f}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,driver_version --format=csv > gpurun_out/${TAG}_gpu.txt
timeout -k 5 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; cat gpurun_out/${TAG}_smoke.log
timeout -k 5 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_C2_m1.json 2> gpurun_out/${TAG}_bench_C2_m1.err; cat gpurun_out/${TAG}_bench_C2_m1.json
timeout -k 5 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cat gpurun_out/${TAG}_bench_reference.json
for spec in "C1 1" "C3 1" "C4 1" "C2 10"; do
  set -- $spec
  timeout -k 5 300 python bench.py --config $1 --m $2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$1_m$2.json 2> gpurun_out/${TAG}_bench_$1_m$2.err
  cut -c1-300 gpurun_out/${TAG}_bench_$1_m$2.json
done
rm -f gpurun_out/${TAG}_sweep_c5.jsonl
for cand in 200 1000 5000; do for part in 20 100; do
  timeout -k 5 300 python bench.py --config C2 --cand $cand --part $part --steps 5 --warmup 3 --no-cpu-baseline >> gpurun_out/${TAG}_sweep_c5.jsonl 2>/dev/null
done; done
wc -l gpurun_out/${TAG}_sweep_c5.jsonl
timeout -k 5 300 python tools/train_bench.py > gpurun_out/${TAG}_train_bench.log 2>&1; cat gpurun_out/${TAG}_train_bench.log
# launch list of the bench command (per-launch times under ncu are serialised and cold: the SHARE is what must agree)
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for spec in "C2 1 rollout_tcs" "C4 1 rollout_tcs" "C2 10 rollout_tc"; do
  set -- $spec
  timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s 6 -c 1 -f -o gpurun_out/${TAG}_$3_$1_m$2 python tools/prof_one.py $1 $2 3 > gpurun_out/${TAG}_ncu_$1_m$2.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_$1_m$2.log
done
timeout -k 5 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_training.py -x -q -m gpu -k "pets and halfcheetah or cadm_steps and ant" > gpurun_out/${TAG}_memcheck_trainer.log 2>&1; tail -5 gpurun_out/${TAG}_memcheck_trainer.log
ls -la gpurun_out | grep ${TAG}
