#!/bin/bash
# tests + bench + launch list + ncu capture of kernels matching $1
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cycle.py -q -x 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print("value %.4e ms/step %.3f frac %.3f stage_ms %s e2e %s launches %s" % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms'], d['e2e'] and d['e2e']['value'], d['gpu_launches']))
PY
tail -3 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum --clock-control none -s 30 -c 12 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
if [ -n "$1" ]; then
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${2:-2} -c ${3:-1} \
    -f -o gpurun_out/${4:-prof} python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_${4:-prof}.log 2>&1
echo "ncu rc=$?"
fi
