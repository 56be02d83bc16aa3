#!/bin/bash
# x1 pass comparison: tile kernel (default) vs chunk kernel (AB200_X1=chunk)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cycle.py tests/test_gpu_baseline_shapes.py tests/test_gpu_tasks.py -m gpu -q -x 2>&1 | tail -4
for x in tile chunk; do
  AB200_X1=$x timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --path three_pass > gpurun_out/bench_x1_$x.json 2> gpurun_out/bench_x1_$x.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_x1_$x.json").read().strip().splitlines()[-1])
    print("x1=$x ms/step %.4f" % d["ms_per_step"], "value %.4g" % d["value"], "roofline", d.get("roofline",{}).get("frac"), d.get("roofline",{}).get("stage_ms"))
except Exception as e:
    print("no line", e); print(open("gpurun_out/bench_x1_$x.err").read()[-1500:])
PY
done
