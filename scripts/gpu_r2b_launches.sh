#!/bin/bash
# launch lists (per-kernel times) of config 5 (task path with flux correction) and config 3 with drag
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 150 -c 260 --csv \
    --log-file gpurun_out/r02b_launches_config5.csv python bench.py --config 5 --steps 1 --warmup 1 > gpurun_out/ncu_cfg5.log 2>&1
echo "cfg5 rc=$?"
python scripts/launch_shares.py gpurun_out/r02b_launches_config5.csv | head -30
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 80 --csv \
    --log-file gpurun_out/r02b_launches_config3_drag.csv python bench.py --config 3 --steps 1 --warmup 1 > gpurun_out/ncu_cfg3.log 2>&1
echo "cfg3 rc=$?"
python scripts/launch_shares.py gpurun_out/r02b_launches_config3_drag.csv | head -30
