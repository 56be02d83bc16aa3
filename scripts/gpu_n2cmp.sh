#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=2
for tag in overlap nooverlap; do
  if [ $tag = nooverlap ]; then export AB200_NO_OVERLAP=1; fi
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e > gpurun_out/cmp_$tag.json 2> gpurun_out/cmp_$tag.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/cmp_$tag.json").read().splitlines()[-1])
print("$tag ms/step %.3f host_issue_ms %.3f halo_exchange_ms %.3f"%(d["ms_per_step"], d["config"]["host_issue_ms_per_step"], d["config"]["halo_exchange_ms"]))
PY
done
