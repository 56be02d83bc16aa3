#!/bin/bash
# final check of the round on one GPU: smoke, the whole GPU suite, the default bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("value %.4e ms/step %.3f frac %.3f traffic %s e2e %.4e cpu %.3e (%s, %d cores) launches %d clocks %s" % (
    d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], d["e2e"]["value"],
    d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"], d["gpu_launches"], d["clocks"]))
PY
tail -3 gpurun_out/bench.err
