#!/bin/bash
# N-rank (default 2) correctness check on real GPUs + weak-scaling bench point
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/tools/check_multigpu.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -12
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 1 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -5
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().splitlines()[-1])
print("N=$N ms/step %.3f value %.4e halo_exchange_ms %s e2e %s"%(d["ms_per_step"], d["value"], d["config"]["halo_exchange_ms"], d["e2e"] and d["e2e"]["value"]))
PY
