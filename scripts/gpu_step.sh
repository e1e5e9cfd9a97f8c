# Step-kernel round: GPU tests, kbench A/B of the cart-pole step kernel, the step / rollout bench lines.
# usage (under gpurun): bash scripts/gpu_step.sh [tag]
TAG=${1:-r01e}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
tools/kbench/kbench_cartpole 1048576 8 400 > gpurun_out/kbench_${TAG}.txt 2>&1; cat gpurun_out/kbench_${TAG}.txt
for n in 4096 32768 262144; do tools/kbench/kbench_cartpole $n 64 400 "(shipped)" | tail -1; done >> gpurun_out/kbench_${TAG}.txt 2>&1
for n in 4194304 16777216 67108864; do tools/kbench/kbench_cartpole $n 2 40 "(shipped)" | tail -1; done >> gpurun_out/kbench_${TAG}.txt 2>&1; tail -6 gpurun_out/kbench_${TAG}.txt
for w in c2 c1 rollout rollout_rec; do
  python bench.py --workload $w > gpurun_out/bench_${TAG}_$w.json 2> gpurun_out/bench_${TAG}_$w.err; echo "bench $w exit $?"; cut -c1-200 gpurun_out/bench_${TAG}_$w.json; tail -3 gpurun_out/bench_${TAG}_$w.err
done
