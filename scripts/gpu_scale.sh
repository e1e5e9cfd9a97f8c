# scaling check as the driver launches it: bash scripts/gpu_scale.sh N [tag] [workloads]   (under gpurun --gpus N)
N=${1:-8}; TAG=${2:-r01}; WL=${3:-"c2 c4 c5"}
for w in $WL; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $w --steps ${STEPS:-20} --warmup 3 \
     > gpurun_out/bench_${TAG}_${w}_n$N.json 2> gpurun_out/bench_${TAG}_${w}_n$N.err; echo "bench $w N=$N exit $?"; cut -c1-260 gpurun_out/bench_${TAG}_${w}_n$N.json; tail -2 gpurun_out/bench_${TAG}_${w}_n$N.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_ref_n$N.json 2> gpurun_out/bench_${TAG}_ref_n$N.err; echo "reference arm exit $?"; cut -c1-260 gpurun_out/bench_${TAG}_ref_n$N.json
