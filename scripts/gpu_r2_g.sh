# charged-ball step shape A/B (dev knob EMEI_CB_SHAPE) + the new rollout workloads
mkdir -p gpurun_out
for sh in 18 26 25 24 44; do
  EMEI_CB_SHAPE=$sh timeout 200 python bench.py --workload c4 --steps 10 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 shape $sh ms/step %.4f frac %.3f e2e %.4g' % (d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))" >> gpurun_out/r2g_c4_shapes.txt
done
EMEI_CB_SHAPE=26 timeout 300 python -m pytest tests -m gpu -q -k "charged or c4_ or step_host" > gpurun_out/r2g_pytest_cb.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest_cb.log
for w in i2p_rollout rollout; do for dt in f32 f64; do
  timeout 300 python bench.py --workload $w --dtype $dt --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_${w}_$dt.json 2> gpurun_out/r2g_${w}_$dt.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/r02_bench_${w}_$dt.json').read()); print('$w $dt value %.4g ms/step %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))" >> gpurun_out/r2g_c4_shapes.txt 2>&1
done; done
cat gpurun_out/r2g_c4_shapes.txt; tail -3 gpurun_out/r2g_pytest_cb.log
