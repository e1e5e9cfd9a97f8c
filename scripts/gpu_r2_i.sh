mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "i2p or I2P or charged or c4_ or rollout_ref or step_host" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 ms/step %.5f frac %.3f value %.4g' % (d['ms_per_step'], d['roofline']['frac'], d['value']))" >> gpurun_out/r2i.txt; }
for dt in f32 f64; do timeout 200 python bench.py --workload i2p --dtype $dt --steps 20 --warmup 5 --no-cpu 2>/dev/null | show "i2p $dt"; done
for dt in f32 f64; do timeout 200 python bench.py --workload c4 --dtype $dt --steps 10 --no-cpu 2>/dev/null | show "c4 $dt"; done
cat gpurun_out/r2i.txt; tail -3 gpurun_out/r2i_pytest.log
