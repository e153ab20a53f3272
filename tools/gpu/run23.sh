export PB200_BACKTRACE=1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_s.json 2> /dev/null || echo "bench failed"
PB200_HOST_THREADS=4 taskset -c 0-3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_s_4thr.json 2> /dev/null || echo "bench failed"
PB200_PROFILE_HOST=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep pb200 | tail -10 > gpurun_out/r02_prof_s.txt
