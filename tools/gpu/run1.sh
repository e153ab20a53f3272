set -x
nvidia-smi --query-gpu=name --format=csv,noheader; nproc
python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_a_par.json 2> gpurun_out/r02_bench_a_par.err
PB200_REPLAY_MODE=seq python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_a_seq.json 2> gpurun_out/r02_bench_a_seq.err
PB200_HOST_THREADS=4 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_a_par_t4.json 2> gpurun_out/r02_bench_a_par_t4.err
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize.py 2>&1 | tail -5
