# Full GPU check of the current tree: whole GPU test suite, smoke, one kept bench line per BASELINE config.
#   gpurun --timeout 1500 -- 'bash tools/r2_full.sh <tag>'
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-r2}
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; cat gpurun_out/${TAG}_smoke.log
for spec in "C2 1" "C1 1" "C3 1" "C4 1" "C2 10"; do
  set -- $spec
  timeout -k 5 300 python bench.py --config $1 --m $2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$1_m$2.json 2> gpurun_out/${TAG}_bench_$1_m$2.err
  cat gpurun_out/${TAG}_bench_$1_m$2.json
done
