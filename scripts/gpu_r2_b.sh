# round 2, second GPU pass: full parity suite, graph pre-sleep sensitivity, flat scoring, ncu captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
for ps in 0 200000 2000000; do for rep in 1 2 3; do
  timeout 200 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu --presleep $ps 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('presleep $ps rep $rep us/step %.3f frac %.3f' % (d['ms_per_step']*1e3, d['roofline']['frac']))" >> gpurun_out/r2b_presleep.txt
done; done
timeout 200 python bench.py --workload c2 --steps 2000 --warmup 5 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('K=2000 us/step %.3f frac %.3f' % (d['ms_per_step']*1e3, d['roofline']['frac']))" >> gpurun_out/r2b_presleep.txt
for w in c3_hopper c3_halfcheetah; do
  timeout 300 python bench.py --workload $w --steps 10 --no-cpu > gpurun_out/r2b_bench_$w.json 2> gpurun_out/r2b_bench_$w.err
done
# steady-state DRAM traffic of the C2 step: 16 consecutive rotated launches, caches NOT flushed between them
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
    -k regex:cartpole_step_f32_tma -s 24 -c 16 --csv --log-file gpurun_out/r02_traffic_c2_range.csv \
    python bench.py --workload c2 --steps 40 --warmup 5 --no-cpu --launch stream > gpurun_out/r2b_ncu_c2_range.log 2>&1
cap() { local w=$1 k=$2 s=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 2 -f -o gpurun_out/prof_r02_$w python bench.py --steps 6 --warmup 3 --no-cpu "$@" > gpurun_out/ncu_$w.log 2>&1
  python scripts/ncu_summary.py gpurun_out/prof_r02_$w.ncu-rep > gpurun_out/r02_ncu_full_$w.txt 2>/dev/null; rm -f gpurun_out/prof_r02_$w.ncu-rep; }
cap c2 cartpole_step_f32_tma 2 --workload c2 --launch stream
cap c2_f64 cartpole_step_kernel 2 --workload c2 --dtype f64 --launch stream
cap c3_hopper_seq reward_terminal_seq 1 --workload c3_hopper_seq
cap c3_halfcheetah_seq reward_terminal_seq 1 --workload c3_halfcheetah_seq
cap c3_hopper "reward_terminal_kernel" 1 --workload c3_hopper
cap sumsq sumsq_kernel 1 --workload c3_hopper_seq
tail -5 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_presleep.txt
