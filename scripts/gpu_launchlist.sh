# ncu launch lists (the "--metrics gpu__time_duration.sum --clock-control none" pass of B200_PROFILING.md, plus instruction
# and DRAM counters) of every bench workload's own command. usage (under gpurun, one GPU): bash scripts/gpu_launchlist.sh [tag] [workloads]
TAG=${1:-r01}
WL=${2:-"c2 c1 c2_large c3_hopper c3_halfcheetah c4 c5 c4_rollout rollout rollout_rec"}
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
for w in $WL; do
  extra=""; [ $w = c4_rollout ] && extra="--total-log2 24"
  python bench.py --workload $w --steps 6 --warmup 3 --no-cpu --launch stream $extra > gpurun_out/plain_$w.log 2>&1 &&
  ncu --metrics $M --clock-control none -c 40 --csv --log-file gpurun_out/launches_${TAG}_$w.csv python bench.py --workload $w --steps 6 --warmup 3 --no-cpu --launch stream $extra > gpurun_out/ncu_$w.log 2>&1
  echo "$w exit $?"; grep -c gpu__time_duration gpurun_out/launches_${TAG}_$w.csv
done
