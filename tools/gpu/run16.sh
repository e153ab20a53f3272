export PB200_BACKTRACE=1
python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py tests/test_zz_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_o.json 2> gpurun_out/r02_bench_o.err || echo "bench failed"
python bench.py --seed 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_o_seed3.json 2> gpurun_out/r02_bench_o_seed3.err || echo "bench failed"
PB200_PROFILE_HOST=1 python bench.py --seed 3 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep pb200 | tail -9 > gpurun_out/r02_prof_o_seed3.txt
