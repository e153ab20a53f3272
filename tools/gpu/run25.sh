export PB200_BACKTRACE=1
python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
python bench.py --seed 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_u_seed3.json 2> /dev/null || echo "bench failed"
PB200_PROFILE_HOST=1 python bench.py --seed 3 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "literal\|lcb ms" | tail -3 > gpurun_out/r02_prof_u_seed3.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_u.json 2> /dev/null || echo "bench failed"
