mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 us/step %.3f frac %.3f e2e %.4g' % (d['ms_per_step']*1e3, d['roofline']['frac'], d['e2e']['value']))" >> gpurun_out/r2n.txt; }
for rep in 1 2 3; do timeout 200 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu 2>/dev/null | show "bench c2 K=20 rep $rep"; done
timeout 200 python bench.py --workload c2 --steps 2000 --warmup 5 --no-cpu 2>/dev/null | show "bench c2 K=2000"
for rep in 1 2; do timeout 200 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu 2>/dev/null | show "bench c1 rep $rep"; done
timeout 200 python bench.py --workload c1 --steps 2000 --warmup 5 --no-cpu 2>/dev/null | show "bench c1 K=2000"
timeout 200 python bench.py --workload c2_large --steps 20 --warmup 5 --no-cpu 2>/dev/null | show "bench c2_large"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
timeout 60 tools/kbench/kbench_trace 1048576 8 400 | grep -v "^cta" | tail -42 > gpurun_out/r2n_trace.txt 2>&1
cat gpurun_out/r2n.txt; tail -3 gpurun_out/r2n_pytest.log; head -3 gpurun_out/r2n_trace.txt
