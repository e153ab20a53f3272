set -x
python -m pytest tests/test_gpu_engine.py -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_b_dev.json 2> gpurun_out/r02_bench_b_dev.err
PB200_REPLAY_MODE=seq python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_b_dev_seq.json 2> gpurun_out/r02_bench_b_dev_seq.err
PB200_HOST_THREADS=4 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_b_dev_t4.json 2> gpurun_out/r02_bench_b_dev_t4.err
PB200_PROFILE_HOST=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep pb200 | tail -30 > gpurun_out/r02_prof_b.txt
python -m pytest tests/test_gpu_fullsize.py tests/test_zz_gpu_fuzz.py tests/test_core_binary.py -m gpu -x -q 2>&1 | tail -5
