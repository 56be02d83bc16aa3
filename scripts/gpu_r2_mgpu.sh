#!/bin/bash
# round 2 multi-GPU visit ($1 = N): real-NCCL bit-identity test, then bench with both transports
cd $GRAFT_REPO_ROOT
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multigpu_nccl.py -m gpu -q -x 2>&1 | tail -15
for t in native torch; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
      bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --no-e2e --transport $t > gpurun_out/bench_n${N}_$t.json 2> gpurun_out/bench_n${N}_$t.err
  echo "bench N=$N $t rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_$t.json").read().strip().splitlines()[-1])
    print("N=$N $t ms/step", d["ms_per_step"], "value %.4g" % d["value"], "halo_ms", d["config"].get("halo_exchange_ms"))
except Exception as e:
    print("no line", e); print(open("gpurun_out/bench_n${N}_$t.err").read()[-2500:])
PY
done
