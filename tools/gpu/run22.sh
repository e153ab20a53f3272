export PB200_BACKTRACE=1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err || echo "bench failed"
python bench.py --seed 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_final_seed3.json 2> /dev/null || echo "bench failed"
python bench.py --workload pop --nq 200 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_final_c3_200q.json 2> /dev/null || echo failed
python bench.py --workload c4 --length 50000000 --nq 64 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_final_c4_50Mbp_64q.json 2> /dev/null || echo failed
PB200_PROFILE_HOST=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep pb200 | tail -10 > gpurun_out/r02_prof_final.txt
PB200_PROFILE_HOST=1 python bench.py --workload c4 --length 50000000 --nq 64 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep pb200 | tail -12 > gpurun_out/r02_prof_final_c4.txt
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_ncu_launches_final.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_l.err
python tools/ncu_traffic.py gpurun_out/r02_ncu_launches_final.csv gpurun_out/ncu_traffic.json > gpurun_out/r02_ncu_traffic_final.txt
ncu --set full --clock-control none --import-source on -k regex:seed_extend_kernel -c 1 -o gpurun_out/r02_ncu_seed python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>> gpurun_out/ncu_r.err
