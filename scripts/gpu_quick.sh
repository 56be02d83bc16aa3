#!/bin/bash
# quick iteration: cycle-level parity tests + bench line (no ncu)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cycle.py -q -x 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print("value %.4e ms/step %.3f frac %.3f stage_ms %s e2e %s launches %s" % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms'], d['e2e'] and d['e2e']['value'], d['gpu_launches']))
PY
tail -5 gpurun_out/bench.err
