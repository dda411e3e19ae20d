# Round checks on the GPU box: the whole GPU test suite, the bench lines (both arms), the ncu launch list and one
# `--set full` capture of the dominant kernel.  Everything lands in gpurun_out/.
set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt
timeout -k 5 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r1_pytest_gpu.log; cat gpurun_out/r1_pytest_gpu.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1; cat gpurun_out/r1_smoke.log
timeout -k 5 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r1_bench_tc3x.json 2> gpurun_out/r1_bench_tc3x.err; cat gpurun_out/r1_bench_tc3x.json
timeout -k 5 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference_cpu.json 2>/dev/null; cat gpurun_out/r1_bench_reference_cpu.json
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 64 --csv --log-file gpurun_out/r1_launches_tc3x.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:rollout_tcs -s 6 -c 1 -f -o gpurun_out/r1_rollout_tcs python tools/prof_one.py C2 1 2 > /dev/null 2>&1
ls -la gpurun_out
