#!/bin/bash
# compute-sanitizer racecheck + memcheck on one small case of the single-pass stage kernel
# (shared-memory slots are reused between phases: the barriers must cover every hazard)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 12 python -m pytest "tests/test_gpu_sweep.py::test_sweep_fast_within_1e12_one_cycle[outflow-bnx0]" -q -x > gpurun_out/racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|hazard" gpurun_out/racecheck.log | sort | uniq -c | head -12
timeout 150 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest "tests/test_gpu_sweep.py::test_sweep_fast_within_1e12_one_cycle[reflect-bnx3]" tests/test_gpu_refine.py::test_partial_variable_ranges_and_one_launch_per_list -q -x > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck.log | head
