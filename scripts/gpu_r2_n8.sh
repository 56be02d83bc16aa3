#!/bin/bash
# N = 8 visit: bit-identity with the peer-write transport, then weak-scaling bench with the
# peer-write transport and with grouped NCCL send/recv, and config 3 strong scaling
cd $GRAFT_REPO_ROOT
N=8
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29558 tests/tools/check_multigpu.py --cycles 3 --transport native 2>&1 | grep check_multigpu | tee gpurun_out/check_multigpu_n8_native_ipc.log
run() {
  env "${@:2}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
      bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --no-e2e > gpurun_out/r02_scale_weak_n8_$1.json 2> gpurun_out/r02_scale_weak_n8_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_scale_weak_n8_$1.json").read().strip().splitlines()[-1])
    print("N=$N $1 ms/step %.4f" % d["ms_per_step"], "value %.4g" % d["value"], "ipc", d["config"].get("peer_write_ipc"), "halo", d["config"].get("halo_exchange_ms"))
except Exception as e:
    print("no line", e); print(open("gpurun_out/r02_scale_weak_n8_$1.err").read()[-1500:])
PY
}
run ipc AB200_DUMMY=1
run nccl AB200_NO_DIRECT=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557 \
  bench.py --gpus $N --config 3 --mesh 512 --no-drag --steps 5 --warmup 3 > gpurun_out/r02_scale_strong_cfg3_n8.json 2> gpurun_out/r02_scale_strong_cfg3_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_scale_strong_cfg3_n8.json').read().strip().splitlines()[-1]); print('cfg3 strong N=8 ms/step %.3f value %.4g' % (d['ms_per_step'], d['value']))"
