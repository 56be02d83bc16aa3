#!/bin/bash
# round 2: interior-only host transfers (e2e A/B), ic boundary tests, march occupancy experiment
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cycle.py -m gpu -q -x -k "interior_only or user_ic or cycles_host" 2>&1 | tail -25 > gpurun_out/pytest_e2e.log
tail -25 gpurun_out/pytest_e2e.log
for x in full interior interior_zc; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 4 --e2e-transfer $x > gpurun_out/bench_e2e_$x.json 2> gpurun_out/bench_e2e_$x.err
  echo "$x rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_e2e_$x.json')); print(d['ms_per_step'], d['roofline']['stage_ms'], d['e2e'])"
done
AB200_VARIANT=mb3 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_mb3.json 2> gpurun_out/bench_mb3.err
echo "mb3 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_mb3.json')); print(d['ms_per_step'], d['roofline']['stage_ms'])"
