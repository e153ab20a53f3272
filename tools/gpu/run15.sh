export PB200_BACKTRACE=1
python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n.json 2> gpurun_out/r02_bench_n.err || echo "bench failed"
python bench.py --workload c4 --length 50000000 --nq 64 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_n_c4.json 2> gpurun_out/r02_bench_n_c4.err || echo failed
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_n.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_l.err
