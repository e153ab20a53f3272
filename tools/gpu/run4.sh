set -x
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err
PB200_HOST_THREADS=8 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_d_t8.json 2> gpurun_out/r02_bench_d_t8.err
PB200_HOST_THREADS=4 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_d_t4.json 2> gpurun_out/r02_bench_d_t4.err
PB200_PROFILE_HOST=1 PB200_REPLAY_DEBUG=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep pb200 | tail -30 > gpurun_out/r02_prof_d.txt
python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
