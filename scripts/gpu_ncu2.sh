#!/bin/bash
# ncu --set full capture: $1 kernel regex, $2 skip, $3 count, $4 output name, $5 bench --path
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${2:-2} -c ${3:-2} \
    -f -o gpurun_out/${4:-prof} python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --path ${5:-auto} > gpurun_out/ncu_${4:-prof}.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_${4:-prof}.log
