# usage: bash scripts/gpu_r2_multi.sh N   -- the driver's launch of bench.py at N ranks + the concurrent pinned-copy probe
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m_topo_n$N.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $T --master-port 29511 scripts/probe_pcie_ranks.py > gpurun_out/r02_pcie_probe_n$N.txt 2> gpurun_out/r2m_probe_n$N.err
( time timeout 400 $T --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_default_n$N.json 2> gpurun_out/r2m_bench_n$N.err ) 2> gpurun_out/r2m_bench_n$N.time
timeout 200 $T --master-port 29513 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n$N.json 2> gpurun_out/r2m_ref_n$N.err
cat gpurun_out/r02_pcie_probe_n$N.txt; cat gpurun_out/r2m_bench_n$N.time; tail -3 gpurun_out/r2m_bench_n$N.err; head -c 600 gpurun_out/r02_bench_default_n$N.json
