#!/bin/bash
# multi-GPU visit ($1 = N): bit-identity on real GPUs with source terms + diffusion configured,
# then config 4 (spherical disk, deck physics) strong-scaled over N GPUs
cd $GRAFT_REPO_ROOT
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multigpu_nccl.py -m gpu -q -x -k "diffusion" 2>&1 | tail -6
cat gpurun_out/check_multigpu_n${N}_native_physics.log | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 \
    bench.py --config 4 --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_cfg4_deck_n${N}.json 2> gpurun_out/bench_cfg4_deck_n${N}.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_cfg4_deck_n${N}.json").read().strip().splitlines()[-1])
    print("cfg4 deck N=$N ms/step %.4f" % d["ms_per_step"], "value %.4g" % d["value"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/bench_cfg4_deck_n${N}.err").read()[-2500:])
PY
