# Rollout round: GPU tests, rollout benches, launch list with instruction counts. usage: bash scripts/gpu_rollout.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for w in rollout rollout_rec c4_rollout; do
  python bench.py --workload $w --steps 5 > gpurun_out/bench_${TAG}_$w.json 2> gpurun_out/bench_${TAG}_$w.err; echo "bench $w exit $?"; cut -c1-600 gpurun_out/bench_${TAG}_$w.json; tail -3 gpurun_out/bench_${TAG}_$w.err
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:rollout -c 4 --csv --log-file gpurun_out/launches_${TAG}_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_$w.log 2>&1
  tail -3 gpurun_out/launches_${TAG}_$w.csv | cut -c1-400
done
