#!/bin/bash
# round 2: GPU parity suite (+ smoke) and a short bench of both stage paths
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -45 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
