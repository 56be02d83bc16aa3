#!/bin/bash
# quick loop: one parity test file subset + bench of the given path(s)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -q -x -k "config2 and three_pass" 2>&1 | tail -3
for p in ${@:-three_pass}; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --path $p > gpurun_out/bench_$p.json 2> gpurun_out/bench_$p.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$p.json").read().strip().splitlines()[-1])
    print("$p", "ms/step %.4f" % d["ms_per_step"], "value %.4g" % d["value"], "roofline", d.get("roofline",{}).get("frac"), d.get("roofline",{}).get("kernel_ms"))
except Exception as e:
    print("no line", e); print(open("gpurun_out/bench_$p.err").read()[-1500:])
PY
done
