mkdir -p gpurun_out
timeout 300 tools/kbench/kbench_cartpole 1048576 8 400 "r2 G" > gpurun_out/r2k_kbench.txt 2>&1
timeout 100 tools/kbench/kbench_cartpole 16777216 2 40 "producer" >> gpurun_out/r2k_kbench.txt 2>&1
timeout 60 tools/kbench/kbench_trace 1048576 8 400 | grep -v "^cta" | tail -45 > gpurun_out/r2k_trace.txt 2>&1
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 us/step %.3f frac %.3f e2e %.4g' % (d['ms_per_step']*1e3, d['roofline']['frac'], d['e2e']['value']))" >> gpurun_out/r2k_kbench.txt; }
for rep in 1 2 3; do timeout 200 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu 2>/dev/null | show "bench c2 K=20 rep $rep"; done
timeout 200 python bench.py --workload c2 --steps 2000 --warmup 5 --no-cpu 2>/dev/null | show "bench c2 K=2000"
timeout 200 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu 2>/dev/null | show "bench c1"
timeout 600 python -m pytest tests -m gpu -q -k "cartpole or step_kernel or mirror or c2_ or c1_protocol or noise or rollout or ip_" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
cat gpurun_out/r2k_kbench.txt; tail -3 gpurun_out/r2k_pytest.log; head -3 gpurun_out/r2k_trace.txt
