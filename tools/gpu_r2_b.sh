#!/bin/bash
mkdir -p gpurun_out
python tools/prof_voxel.py 50e6 2 > gpurun_out/r2b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_pass|k_voxel_reduce|k_voxel_keys16' -s 8 -c 8 -o gpurun_out/r2b_prof python tools/prof_voxel.py 50e6 2 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log
