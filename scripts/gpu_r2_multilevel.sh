#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multilevel.py tests/test_gpu_refine.py -m gpu -q -x 2>&1 | tail -30
