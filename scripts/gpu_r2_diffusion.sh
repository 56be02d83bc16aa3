#!/bin/bash
# round 2: diffusion operators on the GPU + config 4 with the deck's physics
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_diffusion.py tests/test_gpu_sources.py -m gpu -q 2>&1 | tail -12
timeout 600 python bench.py --config 4 --steps 20 --warmup 3 > gpurun_out/bench_cfg4_deck.json 2> gpurun_out/bench_cfg4.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_cfg4_deck.json").read().strip().splitlines()[-1])
print("cfg4 deck ms/step", d["ms_per_step"], "value %.4g" % d["value"], "launches", d["gpu_launches"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --config 4 --steps 2 --warmup 1 > /dev/null 2>&1
