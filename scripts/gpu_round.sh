set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
python bench.py --steps 400 --warmup 5 > gpurun_out/bench1.log 2>&1; echo "bench exit $?"; cat gpurun_out/bench1.log
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu1.log 2>&1
