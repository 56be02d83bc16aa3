#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_cycle.py -m gpu -q -k "user_ic or interior_only" 2>&1 | tail -3
timeout 900 python bench.py --config 5 --steps 3 --warmup 1 > gpurun_out/r02b_bench_config5_smr.json 2> gpurun_out/r02b_cfg5.err
echo "cfg5 rc=$?"; tail -5 gpurun_out/r02b_cfg5.err
timeout 900 python bench.py --config 5 --steps 3 --warmup 1 --no-flux-correction > gpurun_out/r02b_bench_config5_smr_fused.json 2>> gpurun_out/r02b_cfg5.err
echo "cfg5 fused rc=$?"; tail -5 gpurun_out/r02b_cfg5.err
python - <<PY
import json
for n in ("r02b_bench_config5_smr","r02b_bench_config5_smr_fused"):
    try:
        d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        c=d["config"]
        print(n, "ms/step", d["ms_per_step"], "dev", c["device_ms_per_step"], "value %.4g" % d["value"], "exch", c["multilevel_exchange_ms"], "fc", c["flux_correction_ms"], "launches", d["gpu_launches"], c["exchange_descriptors"])
    except Exception as e:
        print(n, "no line", e)
PY
