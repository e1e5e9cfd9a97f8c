mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 us/step %.3f frac %.3f e2e %.4g' % (d['ms_per_step']*1e3, d['roofline']['frac'], d['e2e']['value']))" >> gpurun_out/r2p.txt; }
for rep in 1 2 3; do timeout 200 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu 2>/dev/null | show "bench c2 K=20 rep $rep"; done
timeout 200 python bench.py --workload c2 --steps 200 --warmup 5 --no-cpu 2>/dev/null | show "bench c2 K=200"
timeout 200 python bench.py --workload c2 --steps 2000 --warmup 5 --no-cpu 2>/dev/null | show "bench c2 K=2000"
for rep in 1 2; do timeout 200 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu 2>/dev/null | show "bench c1 rep $rep"; done
timeout 200 python bench.py --workload i2p --steps 20 --warmup 5 --no-cpu 2>/dev/null | show "bench i2p"
cat gpurun_out/r2p.txt
