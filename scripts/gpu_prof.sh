#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_fused_pass -s 6 -c 3 \
    -f -o gpurun_out/fused_prof3 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full3.log 2>&1
echo "ncu rc=$?"
