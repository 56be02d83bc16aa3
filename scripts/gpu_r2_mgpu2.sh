#!/bin/bash
# multi-GPU visit ($1 = N): bit-identity on real GPUs, then bench with and without the
# surface-first overlap (AB200_NO_OVERLAP=1 disables it)
cd $GRAFT_REPO_ROOT
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multigpu_nccl.py -m gpu -q -x -k "native" 2>&1 | tail -6
run() {  # $1 tag, rest: env
  env "${@:2}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
      bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_n${N}_$1.json 2> gpurun_out/bench_n${N}_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_$1.json").read().strip().splitlines()[-1])
    print("N=$N $1 ms/step %.4f" % d["ms_per_step"], "value %.4g" % d["value"], "halo_ms", d["config"].get("halo_exchange_ms"))
except Exception as e:
    print("no line", e); print(open("gpurun_out/bench_n${N}_$1.err").read()[-2500:])
PY
}
run overlap AB200_DUMMY=1
run nooverlap AB200_NO_OVERLAP=1
