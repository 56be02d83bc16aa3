#!/bin/bash
# N-rank weak-scaling bench (one 256^3 tile per GPU over NCCL) + the reference arm for the record.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; cut -c1-700 gpurun_out/bench_n$N.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -20
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "N=1 rc=$?"; cut -c1-300 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err
cut -c1-300 gpurun_out/bench_ref.json
