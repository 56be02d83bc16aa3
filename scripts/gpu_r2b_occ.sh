#!/bin/bash
# march-kernel occupancy variants (96 threads x 224 regs = 9 warps/SM, 160 x 200 = 10 warps/SM)
# against the default (128 x 246 = 8 warps/SM), same box; + the full GPU suite on the default build
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in fast mt96 mt160 fast mt96 mt160; do
  AB200_VARIANT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/occ_$v.json 2> gpurun_out/occ_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/occ_$v.json')); print('$v', d['ms_per_step'], d['roofline']['stage_ms'], d['roofline']['frac'])"
done
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
