#!/bin/bash
# ncu --set full capture of kernels matching $1 (regex), skip $2, count $3 -> gpurun_out/$4.ncu-rep
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${2:-4} -c ${3:-2} \
    -f -o gpurun_out/${4:-prof} python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_${4:-prof}.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_${4:-prof}.log
