export PB200_BACKTRACE=1
for i in 1 2 3 4; do
  python -X faulthandler bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_e$i.json 2> gpurun_out/r02_bench_e$i.err || echo "run $i failed rc $?"
done
grep -l "fatal signal\|Fatal Python" gpurun_out/r02_bench_e*.err
PB200_PROFILE_HOST=1 PB200_REPLAY_DEBUG=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep pb200 | tail -30 > gpurun_out/r02_prof_e.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
