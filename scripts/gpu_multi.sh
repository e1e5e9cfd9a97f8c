# Multi-GPU bench: bash scripts/gpu_multi.sh N [tag]   (under gpurun --gpus N)
N=${1:-2}; TAG=${2:-r01}
python -m pytest tests -m gpu -x -q -k "step_host" > gpurun_out/pytest_gpu_host.log 2>&1; tail -3 gpurun_out/pytest_gpu_host.log
for w in ${WL:-c2 c2_large c4 c4_rollout c3_hopper c5 rollout}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w \
     > gpurun_out/bench_${TAG}_${w}_n$N.json 2> gpurun_out/bench_${TAG}_${w}_n$N.err; echo "bench $w N=$N exit $?"; cut -c1-300 gpurun_out/bench_${TAG}_${w}_n$N.json; tail -2 gpurun_out/bench_${TAG}_${w}_n$N.err
done
python bench.py --workload c2 > gpurun_out/bench_${TAG}_c2_n1b.json 2>/dev/null; cut -c1-200 gpurun_out/bench_${TAG}_c2_n1b.json
