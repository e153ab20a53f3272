set -x
PB200_REPLAY_DEBUG=1 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 >/dev/null | grep "pb200 replay" | sort | uniq -c | sort -rn | head -20 > gpurun_out/r02_dbg_c.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c_dev.json 2> gpurun_out/r02_bench_c_dev.err
PB200_REPLAY_MODE=seq python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c_dev_seq.json 2> gpurun_out/r02_bench_c_dev_seq.err
python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
