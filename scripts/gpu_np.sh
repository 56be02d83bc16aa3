#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for np in 7 5 3 2 1; do
  echo "NP=$np"; AB200_FUSED_NP=$np python bench.py --steps 5 --warmup 2 --no-cpu --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  value %.3e  ms/step %.2f stage_ms %s' % (d['value'], d['ms_per_step'], d['roofline']['stage_ms']))
"
done
ncu --set full --clock-control none --import-source on -k regex:k_fused_pass -s 6 -c 3 \
    -f -o gpurun_out/fused_prof2 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full2.log 2>&1
echo "ncu rc=$?"
