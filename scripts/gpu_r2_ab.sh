#!/bin/bash
# A/B inside ONE box ($1 = N): overlap on + high-priority comm stream | overlap on + default
# priority | overlap off; two repeats each, interleaved
cd $GRAFT_REPO_ROOT
N=${1:-4}
mkdir -p gpurun_out
run() {
  env "${@:2}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
      bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$1.json").read().strip().splitlines()[-1])
    print("N=$N $1 ms/step %.4f" % d["ms_per_step"], "value %.4g" % d["value"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/ab_$1.err").read()[-1500:])
PY
}
for rep in 1 2; do
  run hiprio_$rep AB200_DUMMY=1
  run loprio_$rep AB200_COMM_LOW_PRIO=1
  run nooverlap_$rep AB200_NO_OVERLAP=1
done
