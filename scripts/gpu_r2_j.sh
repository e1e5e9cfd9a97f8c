mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 ms/step %.5f frac %.3f value %.4g' % (d['ms_per_step'], d['roofline']['frac'], d['value']))" >> gpurun_out/r2j.txt; }
for rep in 1 2; do timeout 200 python bench.py --workload i2p --steps 20 --warmup 5 --no-cpu 2>/dev/null | show "i2p f32 minb4 rep $rep"; done
timeout 200 python bench.py --workload i2p_rollout --steps 5 --warmup 3 --no-cpu 2>/dev/null | show "i2p_rollout f32"
timeout 400 python -m pytest tests -m gpu -q -k "i2p or I2P or rollout_ref" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
cat gpurun_out/r2j.txt; tail -3 gpurun_out/r2j_pytest.log
