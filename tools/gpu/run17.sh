export PB200_BACKTRACE=1
PB200_CHECK_SPANS=1 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py tests/test_zz_gpu_fuzz.py tests/test_core_binary.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_p.json 2> gpurun_out/r02_bench_p.err || echo "bench failed"
PB200_PROFILE_HOST=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep pb200 | tail -9 > gpurun_out/r02_prof_p.txt
PB200_HOST_THREADS=4 taskset -c 0-3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_p_4thr.json 2> gpurun_out/r02_bench_p_4thr.err || echo "bench failed"
PB200_PROFILE_HOST=1 PB200_HOST_THREADS=4 taskset -c 0-3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep pb200 | tail -9 > gpurun_out/r02_prof_p_4thr.txt
