#!/bin/bash
# final scaling evidence for one N ($1): config 2 weak scaling (256^3 per GPU) and config 3
# strong scaling (512^3 gas + 4 dust in total, no drag), native transport, overlap on
cd $GRAFT_REPO_ROOT
N=${1:-2}
mkdir -p gpurun_out
launch() { # out-file, args...
  local out=$1; shift
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557 \
      bench.py --gpus $N "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$out.json").read().strip().splitlines()[-1])
    print("$out ms/step %.4f" % d["ms_per_step"], "value %.4g" % d["value"], "scaling", d["scaling"], "halo_ms", d["config"].get("halo_exchange_ms"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
except Exception as e:
    print("$out no line", e); print(open("gpurun_out/$out.err").read()[-2000:])
PY
}
launch r02_scale_weak_n$N --steps 20 --warmup 3 --no-cpu
launch r02_scale_strong_cfg3_n$N --config 3 --mesh 512 --no-drag --steps 5 --warmup 3
