#!/bin/bash
# compute-sanitizer memcheck on this session's new kernels: k_flux_correct / k_box_copy /
# k_block_bcs (multilevel), k_host_rows (zero-copy interior transfers), the AB200_BC_FIXED
# branches of k_fill_ghosts / k_physical_bc, the k_drag fast path
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest \
  "tests/test_gpu_multilevel.py::test_flux_correct_equals_cpu_restriction" \
  "tests/test_gpu_multilevel.py::test_multilevel_task_cycles_strict_bit_identical" \
  "tests/test_gpu_cycle.py::test_cycles_host_interior_only_transfers" \
  "tests/test_gpu_cycle.py::test_user_ic_boundaries_bit_identical" \
  tests/test_gpu_sources.py -k "not history" -q -x > gpurun_out/r02b_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r02b_memcheck.log | sort | uniq -c | head
