# round 2, third GPU pass: full parity suite (new rollout / selector tests), default bench line, i2p, f64 lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2c_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err ) 2> gpurun_out/r2c_bench_default.time
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2c_bench_reference.json 2> gpurun_out/r2c_bench_reference.err
for w in i2p c1; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu > gpurun_out/r2c_bench_$w.json 2> gpurun_out/r2c_bench_$w.err; done
for w in c4 c3_hopper c3_halfcheetah c3_hopper_seq c3_halfcheetah_seq c1; do timeout 300 python bench.py --workload $w --dtype f64 --steps 10 --no-cpu > gpurun_out/r2c_bench_${w}_f64.json 2> gpurun_out/r2c_bench_${w}_f64.err; done
tail -8 gpurun_out/r2c_pytest.log; tail -2 gpurun_out/r2c_smoke.log; cat gpurun_out/r2c_bench_default.time
