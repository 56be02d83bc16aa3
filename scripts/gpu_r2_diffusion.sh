#!/bin/bash
# round 2: diffusion operators on the GPU
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_diffusion.py -m gpu -q 2>&1 | tail -40
