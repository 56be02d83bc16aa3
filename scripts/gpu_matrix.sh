#!/bin/bash
# Stage-path matrix + the GPU suite + both bench paths.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python scripts/stage_matrix.py 2>&1 | tail -12
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
echo "gpu tests:"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cut -c1-200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --path single_pass > gpurun_out/bench_sweep.json 2>> gpurun_out/bench.err
echo "sweep bench:"; cut -c1-200 gpurun_out/bench_sweep.json
python - <<'PY'
import json
for f in ("bench.json", "bench_sweep.json"):
    try:
        d = json.load(open("gpurun_out/" + f))
        print(f, "value %.4e ms/step %.3f frac %.3f stage_ms %s kernel %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["stage_ms"], d["roofline"]["kernel"][:40]))
    except Exception as e:
        print(f, "parse failed", e)
PY
