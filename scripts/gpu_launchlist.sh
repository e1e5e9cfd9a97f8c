# ncu launch lists with instruction counts for the step / rollout workloads. usage (under gpurun): bash scripts/gpu_launchlist.sh [tag]
TAG=${1:-r01e}
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
for w in c2 c1 rollout rollout_rec; do
  python bench.py --workload $w --steps 6 --warmup 3 --no-cpu > gpurun_out/plain_$w.log 2>&1 &&
  ncu --metrics $M --clock-control none -k regex:"cartpole|rollout" -c 8 --csv --log-file gpurun_out/launches_${TAG}_$w.csv python bench.py --workload $w --steps 6 --warmup 3 --no-cpu > gpurun_out/ncu_$w.log 2>&1
  echo "$w exit $?"; tail -8 gpurun_out/launches_${TAG}_$w.csv | cut -c1-300
done
