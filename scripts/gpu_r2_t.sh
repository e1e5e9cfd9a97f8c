# L2 cache hints on the C2 step: evict-first bulk loads (h1), + streaming stores (h2) against the shipped kernel (h0)
mkdir -p gpurun_out
for rep in 1 2; do for h in 0 1 2; do echo "== hints=$h rep $rep" >> gpurun_out/r2t_kbench.txt; timeout 120 tools/kbench/kbench_h$h 1048576 8 400 "r5 G4" >> gpurun_out/r2t_kbench.txt 2>&1; done; done
for h in 0 1 2; do echo "== hints=$h 2^24" >> gpurun_out/r2t_kbench.txt; timeout 120 tools/kbench/kbench_h$h 16777216 2 50 "r5 G4" >> gpurun_out/r2t_kbench.txt 2>&1; done
cat gpurun_out/r2t_kbench.txt
