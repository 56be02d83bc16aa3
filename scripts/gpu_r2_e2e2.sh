#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cycle.py -m gpu -q -k "interior_only or user_ic or cycles_host" 2>&1 | tail -25 > gpurun_out/pytest_e2e.log
tail -25 gpurun_out/pytest_e2e.log
timeout 600 python scripts/probes/e2e_ab.py > gpurun_out/e2e_ab.log 2>&1
tail -12 gpurun_out/e2e_ab.log
