#!/bin/bash
# A/B inside ONE box ($1 = N): peer-write transport (CUDA IPC) vs grouped NCCL send/recv,
# after the real-GPU bit-identity test of both
cd $GRAFT_REPO_ROOT
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multigpu_nccl.py -m gpu -q -x -k "native" 2>&1 | tail -4
cat gpurun_out/check_multigpu_n${N}_native.log | head -3
AB200_NO_DIRECT=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29558 tests/tools/check_multigpu.py --cycles 3 --transport native 2>&1 | grep check_multigpu
run() {
  env "${@:2}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
      bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$1.json").read().strip().splitlines()[-1])
    print("N=$N $1 ms/step %.4f" % d["ms_per_step"], "value %.4g" % d["value"], "ipc", d["config"].get("peer_write_ipc"))
except Exception as e:
    print("no line", e); print(open("gpurun_out/ab_$1.err").read()[-1500:])
PY
}
for rep in 1 2; do
  run direct_$rep AB200_DUMMY=1
  run nccl_$rep AB200_NO_DIRECT=1
done
