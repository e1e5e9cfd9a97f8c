# round 2, consolidated pass: parity suite, smoke, default bench line + reference arm, launch lists, remaining ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2e_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r2e_bench_default.err ) 2> gpurun_out/r2e_bench_default.time
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r2e_bench_reference.err
timeout 300 python bench.py --workload c2 --steps 2000 --warmup 5 --no-cpu > gpurun_out/r02_bench_c2_k2000.json 2>/dev/null
for w in c2_large i2p rollout rollout_rec c5_seq; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_bench_$w.json 2> gpurun_out/r2e_bench_$w.err; done
timeout 200 tools/kbench/kbench_cartpole 1048576 8 400 "producer = group 3" > gpurun_out/r2e_kbench.txt 2>&1
KBENCH_REPS=1 timeout 200 tools/kbench/kbench_cartpole 1048576 8 400 "producer = group 3" >> gpurun_out/r2e_kbench.txt 2>&1
bash scripts/gpu_launchlist.sh r02 "c2 c3_hopper_seq c3_halfcheetah_seq c5_seq c3_hopper c4" > gpurun_out/r2e_launchlist.log 2>&1
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -c 40 --csv --log-file gpurun_out/launches_r02_c2_f64.csv python bench.py --workload c2 --dtype f64 --steps 6 --warmup 3 --no-cpu --launch stream > gpurun_out/ncu_c2_f64.log 2>&1
tail -6 gpurun_out/r2e_pytest.log; tail -2 gpurun_out/r2e_smoke.log; cat gpurun_out/r2e_bench_default.time; cat gpurun_out/r2e_kbench.txt
