# rollout-only round: the rollout tests, benches and a launch list with stall metrics. usage: bash scripts/gpu_rollout_quick.sh [tag]
TAG=${1:-r01j}
python -m pytest tests -m gpu -x -q -k "rollout or dataset" > gpurun_out/pytest_rollout.log 2>&1; tail -3 gpurun_out/pytest_rollout.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct,smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct,smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum
for w in rollout rollout_rec; do
  python bench.py --workload $w --no-cpu > gpurun_out/bench_${TAG}_$w.json 2> gpurun_out/bench_${TAG}_$w.err; cut -c1-160 gpurun_out/bench_${TAG}_$w.json
  ncu --metrics $M --clock-control none -k regex:rollout -c 3 --csv --log-file gpurun_out/launches_${TAG}_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_$w.log 2>&1
done
