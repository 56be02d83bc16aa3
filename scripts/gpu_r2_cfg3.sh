#!/bin/bash
# config 3 (gas + 4 dust, PLM+HLLE, periodic) strong-scaling bench: $1 = N GPUs, $2 = mesh
cd $GRAFT_REPO_ROOT
N=${1:-1}; M=${2:-512}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --config 3 --mesh $M --steps 5 --warmup 3 > gpurun_out/bench_cfg3_n1_m$M.json 2> gpurun_out/bench_cfg3_n1_m$M.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 \
      bench.py --config 3 --mesh $M --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_cfg3_n${N}_m$M.json 2> gpurun_out/bench_cfg3_n${N}_m$M.err
fi
echo "rc=$?"; cut -c1-600 gpurun_out/bench_cfg3_n${N}_m$M.json; tail -5 gpurun_out/bench_cfg3_n${N}_m$M.err; free -g | head -2
