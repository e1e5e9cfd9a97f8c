# round 2, final single-GPU pass: parity suite, smoke, default line + reference arm (+ optional extras: EXTRA=1)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2f_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r2f_bench_default.err ) 2> gpurun_out/r2f_bench_default.time
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r2f_bench_reference.err
if [ -n "$EXTRA" ]; then
  for w in c1 rollout rollout_rec c4_rollout i2p; do timeout 300 python bench.py --workload $w --no-cpu > gpurun_out/r02_bench_$w.json 2> gpurun_out/r2f_bench_$w.err; done
  WORKLOADS="rollout c4_rollout" bash scripts/gpu_ncu_full.sh r02 > gpurun_out/r2f_ncu_full.log 2>&1
  bash scripts/gpu_launchlist.sh r02 "rollout_rec" > gpurun_out/r2f_launchlist.log 2>&1
fi
tail -3 gpurun_out/r2f_pytest.log; tail -2 gpurun_out/r2f_smoke.log; cat gpurun_out/r2f_bench_default.time
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default_n1.json'))
print('top %.1f G  %.3f us  frac %.3f e2e %.3f G'%(d['value']/1e9,d['ms_per_step']*1e3,d['roofline']['frac'],d['e2e']['value']/1e9))
for k,v in d['secondary'].items(): print('  ',k,'%.4g'%v['value'], v['roofline']['bound'], '%.3f'%v['roofline']['frac'], 'e2e %.4g'%((v.get('e2e') or {}).get('value') or 0))
r=json.load(open('gpurun_out/r02_bench_reference_n1.json')); print('ref %.1f M'%(r['value']/1e6), r['cpu_baseline']['cores'])
PY
