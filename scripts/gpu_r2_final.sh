# final single-GPU pass of the round: what the driver runs (tests, smoke, default bench line, reference arm)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2final_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r2final_bench.err ) 2> gpurun_out/r2final_bench.time
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r2final_ref.err
for dt in f32 f64; do timeout 300 python bench.py --workload c4 --dtype $dt --steps 10 --no-cpu > gpurun_out/r02_bench_c4_$dt.json 2>/dev/null; done
timeout 300 python bench.py --workload i2p --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_i2p.json 2>/dev/null
tail -3 gpurun_out/r2final_pytest.log; tail -1 gpurun_out/r2final_smoke.log; cat gpurun_out/r2final_bench.time
