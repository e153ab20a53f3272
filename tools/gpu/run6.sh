export PB200_BACKTRACE=1
for k in 1 2; do
  PB200_DISCOVERY_SLICES=$k python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_f$k.json 2> gpurun_out/r02_bench_f$k.err || echo "run $k failed rc $?"
done
PB200_PROFILE_HOST=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep pb200 | tail -30 > gpurun_out/r02_prof_f.txt
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_engine.py tests/test_zz_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
