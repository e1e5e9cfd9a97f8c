mkdir -p gpurun_out
for sh in 44 45 43 83 82 44; do
  EMEI_CB_SHAPE=$sh timeout 200 python bench.py --workload c4 --steps 10 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 shape $sh ms/step %.4f frac %.3f e2e %.4g' % (d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))" >> gpurun_out/r2h_c4_shapes.txt
done
for sh in 18 44; do EMEI_CB_SHAPE=$sh timeout 200 python bench.py --workload c4 --total-log2 23 --steps 20 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 2^23 shape $sh ms/step %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))" >> gpurun_out/r2h_c4_shapes.txt; done
cat gpurun_out/r2h_c4_shapes.txt
