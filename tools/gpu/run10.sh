export PB200_BACKTRACE=1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --workload c4 --length 50000000 --nq 64 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_j_c4.json 2> gpurun_out/r02_bench_j_c4.err || echo failed
