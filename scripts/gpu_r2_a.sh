# round 2, first GPU pass: parity suite, C2 kernel shape variants (kbench), scoring layouts, launch modes, full default line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -x > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 300 tools/kbench/kbench_cartpole 1048576 8 400 r2 > gpurun_out/r2a_kbench.txt 2>&1
timeout 120 tools/kbench/kbench_cartpole 1048576 8 400 stream >> gpurun_out/r2a_kbench.txt 2>&1
for w in c3_hopper c3_hopper_seq c3_halfcheetah c3_halfcheetah_seq; do
  timeout 300 python bench.py --workload $w --steps 10 --no-cpu > gpurun_out/r2a_bench_$w.json 2> gpurun_out/r2a_bench_$w.err
done
timeout 200 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu --launch graph > gpurun_out/r2a_bench_c2_graph.json 2> gpurun_out/r2a_bench_c2_graph.err
timeout 200 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu --launch stream > gpurun_out/r2a_bench_c2_stream.json 2> gpurun_out/r2a_bench_c2_stream.err
timeout 200 python bench.py --workload c2 --dtype f64 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2a_bench_c2_f64.json 2> gpurun_out/r2a_bench_c2_f64.err
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err ) 2> gpurun_out/r2a_bench_default.time
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_kbench.txt; cat gpurun_out/r2a_bench_default.time
