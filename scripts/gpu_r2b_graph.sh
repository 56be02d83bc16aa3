#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cycle.py -m gpu -q -x -k "graph" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -q -k "linwave" 2>&1 | tail -5
timeout 600 python scripts/probes/graph_ab.py > gpurun_out/r02b_graph_ab.log 2>&1
tail -3 gpurun_out/r02b_graph_ab.log | cut -c1-300
