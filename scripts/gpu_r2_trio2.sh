#!/bin/bash
# trio quick loop: parity subset, bench role_split, ncu full of k_trio
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sweep.py -m gpu -q -x -k "role_split and (strict or bnx0)" 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --path role_split > gpurun_out/bench_role_split.json 2> gpurun_out/bench_role_split.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_role_split.json").read().strip().splitlines()[-1])
    print("role_split ms/step", d["ms_per_step"], "value", d["value"], "roofline", d.get("roofline",{}).get("frac"))
except Exception as e:
    print("no line", e); print(open("gpurun_out/bench_role_split.err").read()[-1500:])
PY
if [ "$1" = "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trio -s 2 -c 2 \
    -f -o gpurun_out/${2:-trio_prof} python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --path role_split > gpurun_out/ncu_trio.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_trio.log
fi
