#!/bin/bash
# round 2: warp-specialised single-pass kernel (trio.cuh): parity first (short timeouts: a
# barrier-protocol bug would hang), then stage timings of all paths
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sweep.py -m gpu -q -x -k "role_split" 2>&1 | tail -25 > gpurun_out/pytest_trio.log
echo "trio tests rc=${PIPESTATUS[0]}"; tail -25 gpurun_out/pytest_trio.log
for p in role_split single_pass three_pass; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --path $p > gpurun_out/bench_$p.json 2> gpurun_out/bench_$p.err
  echo "bench $p rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$p.json").read().strip().splitlines()[-1])
    print("$p", "ms/step", d["ms_per_step"], "value", d["value"], "roofline", d.get("roofline",{}).get("frac"), "stage_ms", d.get("config",{}).get("stage_ms"))
except Exception as e:
    print("no line", e); print(open("gpurun_out/bench_$p.err").read()[-1500:])
PY
done
