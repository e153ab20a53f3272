export PB200_BACKTRACE=1
python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err || echo "bench failed"
PB200_NO_DEVICE_ANCHORS=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_i_noanc.json 2> gpurun_out/r02_bench_i_noanc.err || echo "bench failed"
PB200_PROFILE_HOST=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep pb200 | tail -12 > gpurun_out/r02_prof_i.txt
python bench.py --workload pop --nq 200 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_i_c3.json 2> gpurun_out/r02_bench_i_c3.err || echo failed
python -m pytest tests/test_zz_gpu_fuzz.py tests/test_core_binary.py -m gpu -x -q 2>&1 | tail -4
