#!/bin/bash
# weak-scaling point N (default 8) on one box (256^3 per GPU), device-timed value only
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "N=$N rc=$?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -5
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().splitlines()[-1])
print("ms/step %.3f value %.4e halo_exchange_ms %s host_issue_ms %s"%(d["ms_per_step"], d["value"], d["config"]["halo_exchange_ms"], d["config"]["host_issue_ms_per_step"]))
PY
