#!/bin/bash
# multi-rank: loopback parity test on GPU 0, then the N-rank bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multirank.py -q -x 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; cat gpurun_out/bench_n$N.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -20
