#!/bin/bash
# round 2: point-mass gravity + rotating frame (mass-flux tap) on the GPU, then the default bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sources.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_tap.json 2> gpurun_out/bench_tap.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_tap.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value %.4g" % d["value"], "roofline", d["roofline"].get("frac"))
PY
timeout 600 python bench.py --config 4 --steps 20 --warmup 3 > gpurun_out/bench_cfg4.json 2>> gpurun_out/bench_tap.err
cut -c1-330 gpurun_out/bench_cfg4.json
