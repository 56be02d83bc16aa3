#!/bin/bash
# round 2, second session: full GPU suite + smoke, default bench line (interior-only e2e
# transfers), reference arm, config 4 with the deck's ic user BCs, ncu launch list
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r02b_gpu_box.txt
nproc >> gpurun_out/r02b_gpu_box.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/r02b_smoke.log
timeout 2400 python -m pytest tests -m gpu -q --durations=10 2>&1 | tail -40 > gpurun_out/r02b_pytest_gpu.log
tail -14 gpurun_out/r02b_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02b_bench_reference.json 2>> gpurun_out/r02b_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-transfer full > gpurun_out/r02b_bench_e2e_full.json 2>> gpurun_out/r02b_bench.err
timeout 600 python bench.py --config 4 --steps 5 --warmup 2 --no-cpu --no-e2e > gpurun_out/r02b_bench_config4_ic_bcs.json 2>> gpurun_out/r02b_bench.err
echo "cfg4 rc=$?"
python - <<PY
import json
for n in ("r02b_bench","r02b_bench_reference","r02b_bench_e2e_full","r02b_bench_config4_ic_bcs"):
    try:
        d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", d.get("ms_per_step"), "value %.4g" % d["value"], "frac", (d.get("roofline") or {}).get("frac"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
    except Exception as e:
        print(n, "no line", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 140 --csv \
    --log-file gpurun_out/r02b_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02b_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
