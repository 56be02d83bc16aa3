#!/bin/bash
# final-build check at N = 2: real-NCCL bit-identity test, the driver's own bench launch (with the
# e2e leg), the reference arm under torchrun, config 4 at N = 2 (outflow instead of ic BCs)
cd $GRAFT_REPO_ROOT
N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multigpu_nccl.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err
echo "bench N=2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02b_bench_reference_n2.json 2>> gpurun_out/r02b_bench_n2.err
echo "reference N=2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557 \
    bench.py --gpus $N --config 4 --steps 5 --warmup 2 --no-cpu --no-e2e > gpurun_out/r02b_bench_config4_n2.json 2>> gpurun_out/r02b_bench_n2.err
echo "cfg4 N=2 rc=$?"
python - <<PY
import json
for n in ("r02b_bench_n2","r02b_bench_reference_n2","r02b_bench_config4_n2"):
    try:
        d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", d.get("ms_per_step"), "value %.4g" % d["value"], "e2e", (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("d2h_bytes_per_step"))
    except Exception as e:
        print(n, "no line", e); print(open("gpurun_out/r02b_bench_n2.err").read()[-1500:])
PY
