#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_diffusion_flux|k_diffusion_update|k_rotating_frame|k_point_mass|k_finish_stage" -c 5 -o gpurun_out/r02_diffusion python bench.py --config 4 --steps 1 --warmup 1 > /dev/null 2> gpurun_out/ncu_diff.err
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/ncu_diff.err
