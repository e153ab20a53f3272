export PB200_BACKTRACE=1
python -m pytest tests/test_gpu_fullsize.py tests/test_zz_gpu_fuzz.py -m gpu -x -q -k "full_size or final_gaps" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_v.json 2> /dev/null || echo "bench failed"
PB200_HOST_THREADS=4 taskset -c 0-3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_v_4thr.json 2> /dev/null || echo "bench failed"
