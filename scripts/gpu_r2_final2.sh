# final single-GPU pass after the prefetch kernel: tests, smoke, default line, reference arm, C2 captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2final_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r2final_bench.err ) 2> gpurun_out/r2final_bench.time
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r2final_ref.err
timeout 300 python bench.py --workload c2 --steps 2000 --warmup 5 --no-cpu > gpurun_out/r02_bench_c2_k2000.json 2>/dev/null
timeout 300 python bench.py --workload c2_large --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_c2_large.json 2>/dev/null
timeout 300 python bench.py --workload c1 --steps 2000 --warmup 5 --no-cpu > gpurun_out/r02_bench_c1_k2000.json 2>/dev/null
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
    -k regex:cartpole_step_f32_tma -s 24 -c 16 --csv --log-file gpurun_out/r02_traffic_c2_range.csv \
    python bench.py --workload c2 --steps 40 --warmup 5 --no-cpu --launch stream > gpurun_out/r2final_ncu_range.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cartpole_step_f32_tma -s 2 -c 2 -f -o gpurun_out/prof_r02_c2 python bench.py --workload c2 --steps 6 --warmup 3 --no-cpu --launch stream > gpurun_out/ncu_c2.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_r02_c2.ncu-rep > gpurun_out/r02_ncu_full_c2.txt 2>/dev/null; rm -f gpurun_out/prof_r02_c2.ncu-rep
bash scripts/gpu_launchlist.sh r02 "c2 c1 c4" > gpurun_out/r2final_launchlist.log 2>&1
tail -3 gpurun_out/r2final_pytest.log; tail -1 gpurun_out/r2final_smoke.log; cat gpurun_out/r2final_bench.time
