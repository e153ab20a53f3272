nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err || echo failed
PB200_PROFILE_HOST=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 3 --warmup 2 2>&1 >/dev/null | grep "pb200" | tail -40 > gpurun_out/r02_prof_n4.txt
