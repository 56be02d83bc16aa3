#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=4
run() { tag=$1; shift
env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/tune_$tag.json 2> gpurun_out/tune_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/tune_$tag.json").read().splitlines()[-1])
print("$tag", "ms/step %.3f"%d["ms_per_step"], "halo_exchange_ms %.3f"%d["config"]["halo_exchange_ms"], "bytes", d["config"]["halo_bytes_per_exchange"])
PY
}
run base A=1
run ch16 NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32
run ch32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
# (NCCL_P2P_USE_CUDA_MEMCPY=1 hangs the 4-rank job on this box: do not retry)
