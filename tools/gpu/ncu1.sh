set -x
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_l.err
ncu --set full --clock-control none --import-source on -k regex:recursion_level_kernel -c 2 -o gpurun_out/r02_prof_rec python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_r.err
ls -la gpurun_out/*.ncu-rep
