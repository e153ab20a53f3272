export PB200_BACKTRACE=1
python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_k.json 2> gpurun_out/r02_bench_k.err || echo "bench failed"
PB200_PROFILE_HOST=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep pb200 | tail -12 > gpurun_out/r02_prof_k.txt
nproc; lscpu | grep -E 'Model name|Socket|NUMA|^CPU\(s\)'
