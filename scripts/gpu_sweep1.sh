#!/bin/bash
# GPU visit for the single-pass sweep kernel: parity tests (sanitizer on one small case only if
# they fail), bench, launch list and one full ncu capture, then the rest of the GPU suite.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests/test_gpu_sweep.py -q 2>&1 | tail -40 > gpurun_out/pytest_sweep.log
echo "sweep tests:"; tail -25 gpurun_out/pytest_sweep.log
if ! grep -q " passed" gpurun_out/pytest_sweep.log || grep -q "failed" gpurun_out/pytest_sweep.log; then
  timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest "tests/test_gpu_sweep.py::test_sweep_fast_within_1e12_one_cycle[periodic-bnx1]" -q -x 2>&1 | tail -40 > gpurun_out/sanitizer.log
  echo "sanitizer:"; tail -12 gpurun_out/sanitizer.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print("value %.4e ms/step %.3f frac %.3f stage_ms %s e2e %s launches %s" % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms'], d['e2e'] and d['e2e']['value'], d['gpu_launches']))
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/bench.err
AB200_NO_SWEEP=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_3pass.json 2>> gpurun_out/bench.err
echo "3-pass bench:"; cut -c1-400 gpurun_out/bench_3pass.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum --clock-control none -s 20 -c 16 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 2 -c 2 \
    -f -o gpurun_out/sweep_prof python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_sweep.log 2>&1
echo "ncu full rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_sweep.py 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "all gpu tests:"; tail -15 gpurun_out/pytest_gpu.log
ls -la gpurun_out
