#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum --clock-control none -s 30 -c 40 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
echo "rc=$?"
