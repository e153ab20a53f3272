export PB200_BACKTRACE=1
python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py tests/test_zz_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_r.json 2> /dev/null || echo "bench failed"
python bench.py --workload c4 --length 50000000 --nq 64 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_r_c4.json 2> /dev/null || echo failed
python bench.py --workload pop --nq 200 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_r_c3.json 2> /dev/null || echo failed
