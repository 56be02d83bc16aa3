#!/bin/bash
# One GPU-box visit collecting the round's evidence: smoke, GPU parity tests, both bench arms,
# ncu launch list of the bench command, ncu --set full of the stage kernels (both paths).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> gpurun_out/gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cut -c1-250 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cut -c1-250 gpurun_out/bench_ref.json
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --path single_pass > gpurun_out/bench_single_pass.json 2>> gpurun_out/bench.err
cut -c1-250 gpurun_out/bench_single_pass.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e \
    > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_xchunk|k_march" -s 6 -c 3 \
    -f -o gpurun_out/passes_prof python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e \
    > gpurun_out/ncu_full.log 2>&1
echo "ncu full (passes) rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 2 -c 2 \
    -f -o gpurun_out/sweep_prof python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --path single_pass \
    > gpurun_out/ncu_sweep.log 2>&1
echo "ncu full (sweep) rc=$?"
ls -la gpurun_out
