#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sources.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python bench.py --config 3 --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02b_bench_config3_drag.json 2> gpurun_out/r02b_cfg3.err
echo "cfg3 rc=$?"; tail -3 gpurun_out/r02b_cfg3.err
python -c "
import json; d=json.load(open('gpurun_out/r02b_bench_config3_drag.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'])"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_drag -c 2 --csv \
    --log-file gpurun_out/r02b_drag_kernel.csv python bench.py --config 3 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_drag.log 2>&1
python scripts/launch_shares.py gpurun_out/r02b_drag_kernel.csv | head
