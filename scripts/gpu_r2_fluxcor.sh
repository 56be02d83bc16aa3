#!/bin/bash
# round 2: flux correction + multilevel GPU tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multilevel.py -m gpu -q -x --durations=5 2>&1 | tail -30 > gpurun_out/pytest_fluxcor.log
tail -30 gpurun_out/pytest_fluxcor.log
