set -x
cd $GRAFT_REPO_ROOT
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -x -q -k "cem or sharding or kat or ties" 2>&1 | tail -5
timeout -k 5 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
timeout -k 5 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config C3
timeout -k 5 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config C4
timeout -k 5 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config C1
timeout -k 5 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config C2 --m 10
