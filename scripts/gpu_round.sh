# One GPU round: tests, smoke, every bench workload (N=1), launch list of the default bench.
# usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
for w in c2 c1 c2_large i2p c3_hopper c3_halfcheetah c4 c5 c4_rollout rollout rollout_rec; do
  python bench.py --workload $w > gpurun_out/bench_${TAG}_$w.json 2> gpurun_out/bench_${TAG}_$w.err; echo "bench $w exit $?"; cut -c1-400 gpurun_out/bench_${TAG}_$w.json; tail -3 gpurun_out/bench_${TAG}_$w.err
done
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu1.log 2>&1
