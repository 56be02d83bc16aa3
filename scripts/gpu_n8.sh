#!/bin/bash
# weak-scaling point N (default 8) on one box (256^3 per GPU)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 1 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "N=$N rc=$?"; tail -1 gpurun_out/bench_n$N.json | cut -c1-330; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -5
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().splitlines()[-1])
print("ms/step %.3f value %.4e halo_exchange_ms %s bytes %s e2e %s"%(d["ms_per_step"], d["value"], d["config"]["halo_exchange_ms"], d["config"]["halo_bytes_per_exchange"], d["e2e"] and d["e2e"]["value"]))
PY
