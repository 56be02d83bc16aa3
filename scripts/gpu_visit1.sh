#!/bin/bash
# smoke + gpu tests + bench + launch list w/ metrics + ncu full of the three stage kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> gpurun_out/gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 10 -c 40 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_march|k_xchunk" -s 3 -c 3 \
    -f -o gpurun_out/stage_prof python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
