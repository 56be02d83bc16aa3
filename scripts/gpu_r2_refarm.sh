#!/bin/bash
# round 2: default bench line + the reference arm on the SAME 256^3 mesh (no test suite)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2>&1 | grep real
python - <<PY
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value %.4g" % d["value"], "roofline", d["roofline"].get("frac"), "e2e", d["e2e"], "cpu", d["cpu_baseline"])
PY
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err ) 2>&1 | grep real
cut -c1-300 gpurun_out/bench_ref.json
