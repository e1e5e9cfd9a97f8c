# rollout kernels after the policy shift register / spare reset samples: tests, then the three rollout workloads
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "rollout or dataset or charged" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 value %.4g ms/step %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))" >> gpurun_out/r2q.txt; }
for w in rollout rollout_rec c4_rollout; do
  timeout 300 python bench.py --workload $w --no-cpu 2> gpurun_out/r2q_$w.err | tee gpurun_out/r2q_bench_$w.json | show "$w"
done
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for w in rollout c4_rollout; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:rollout -c 2 --csv --log-file gpurun_out/r2q_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu > gpurun_out/r2q_ncu_$w.log 2>&1
done
cat gpurun_out/r2q.txt; tail -3 gpurun_out/r2q_pytest.log
