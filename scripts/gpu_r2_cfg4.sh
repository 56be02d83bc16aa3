#!/bin/bash
# round 2: config 4 (spherical disk) with the deck's sources + alpha viscosity, and without
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --config 4 --steps 20 --warmup 3 > gpurun_out/bench_cfg4_deck.json 2> gpurun_out/bench_cfg4.err
timeout 600 python bench.py --config 4 --no-sources --steps 20 --warmup 3 > gpurun_out/bench_cfg4_nosrc.json 2>> gpurun_out/bench_cfg4.err
python - <<PY
import json
for f in ("bench_cfg4_deck","bench_cfg4_nosrc"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, "ms/step", d["ms_per_step"], "value %.4g" % d["value"], "launches", d["gpu_launches"])
PY
tail -3 gpurun_out/bench_cfg4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --config 4 --steps 2 --warmup 1 > /dev/null 2>&1
python scripts/launch_shares.py gpurun_out/launches_cfg4.csv 2>/dev/null | head -30
