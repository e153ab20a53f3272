export PB200_BACKTRACE=1
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_g_n$N.json 2> gpurun_out/r02_bench_g_n$N.err || echo "failed rc $?"
tail -3 gpurun_out/r02_bench_g_n$N.err
nproc
