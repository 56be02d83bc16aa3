#!/bin/bash
# round-2 evidence on one B200: bench lines (default, shocked state, all stage paths), the
# reference arm, the ncu launch list of the bench command, ncu --set full of the default-path
# stage kernels and of the two single-pass kernels.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_box.txt
nproc >> gpurun_out/gpu_box.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> gpurun_out/gpu_box.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --state shocked > gpurun_out/r02_bench_shocked_state.json 2>> gpurun_out/r02_bench.err
for p in single_pass role_split; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --path $p > gpurun_out/r02_bench_$p.json 2>> gpurun_out/r02_bench.err
done
python - <<PY
import json
for n in ("r02_bench","r02_bench_reference","r02_bench_shocked_state","r02_bench_single_pass","r02_bench_role_split"):
    try:
        d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", d.get("ms_per_step"), "value %.4g" % d["value"], "frac", (d.get("roofline") or {}).get("frac"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
    except Exception as e:
        print(n, "no line", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 140 --csv \
    --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_xchunk|k_march" -s 9 -c 3 \
    -f -o gpurun_out/r02_passes_prof python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
echo "ncu full (passes) rc=$?"
ls -la gpurun_out | tail -15
