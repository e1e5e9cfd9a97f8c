mkdir -p gpurun_out
timeout 300 tools/kbench/kbench_cartpole 1048576 8 400 "r2 G4 direct" > gpurun_out/r2d_kbench.txt 2>&1
timeout 300 tools/kbench/kbench_cartpole 1048576 8 400 "r2 G4 direct" >> gpurun_out/r2d_kbench.txt 2>&1
timeout 200 tools/kbench/kbench_cartpole 16777216 2 40 "r2 G4 direct stores, producer = group 3" >> gpurun_out/r2d_kbench.txt 2>&1
for rep in 1 2 3; do
  timeout 200 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench c2 K=20 rep $rep us/step %.3f frac %.3f' % (d['ms_per_step']*1e3, d['roofline']['frac']))" >> gpurun_out/r2d_kbench.txt
done
timeout 200 python bench.py --workload c2 --steps 2000 --warmup 5 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench c2 K=2000 us/step %.3f frac %.3f' % (d['ms_per_step']*1e3, d['roofline']['frac']))" >> gpurun_out/r2d_kbench.txt
timeout 200 python bench.py --workload c2_large --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench c2_large us/step %.3f frac %.3f' % (d['ms_per_step']*1e3, d['roofline']['frac']))" >> gpurun_out/r2d_kbench.txt
timeout 600 python -m pytest tests -m gpu -q --maxfail=40 -k "cartpole or step_kernel or mirror or c2_ or c1_protocol or noise or rollout" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
cat gpurun_out/r2d_kbench.txt; tail -5 gpurun_out/r2d_pytest.log
