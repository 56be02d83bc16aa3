#!/bin/bash
# round 2: all source terms (incl. the full drag) + diffusion on the GPU
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sources.py tests/test_gpu_diffusion.py -m gpu -q 2>&1 | tail -25
