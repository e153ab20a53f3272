export PB200_BACKTRACE=1
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -3
python bench.py --seed 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_t_seed3.json 2> /dev/null || echo "bench failed"
python bench.py --seed 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_t_seed4.json 2> /dev/null || echo "bench failed"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_t.json 2> /dev/null || echo "bench failed"
PB200_PROFILE_HOST=1 python bench.py --seed 3 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "literal\|lcb ms" | tail -4 > gpurun_out/r02_prof_t_seed3.txt
