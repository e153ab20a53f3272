export PB200_BACKTRACE=1
python bench.py --workload pop --nq 200 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_h_c3.json 2> gpurun_out/r02_bench_h_c3.err || echo failed
PB200_PROFILE_HOST=1 python bench.py --workload pop --nq 200 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep pb200 | tail -20 > gpurun_out/r02_prof_h_c3.txt
python bench.py --workload c4 --length 50000000 --nq 64 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_h_c4.json 2> gpurun_out/r02_bench_h_c4.err || echo failed
python -m pytest tests/test_core_binary.py -m gpu -x -q 2>&1 | tail -3
