#!/bin/bash
# Runs on the GPU box: bench line, ncu launch list, one --set full capture of the fused passes.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref.json
# launch list (cold-cache, serialised): compare SHARES
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e \
    > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
# full capture of the three fused passes of one stage
ncu --set full --clock-control none --import-source on -k regex:k_fused_pass -s 6 -c 3 \
    -f -o gpurun_out/fused_prof python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e \
    > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
