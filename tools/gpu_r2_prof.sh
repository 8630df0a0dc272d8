#!/bin/bash
# round-2 profiles: launch list of one default step, --set full rows of the heavy kernels, the grid-ground kernels, k_obb
mkdir -p gpurun_out
CMD="python bench.py --points 100e6 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-modes"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu1.log 2>&1
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_pass|k_voxel_reduce|k_voxel_keys16|k_chunk_minmax|k_sum_tables|k_compact_xyz|k_db_union|k_db_core2|k_db_cells|k_db_labels_core' -s 22 -c 22 -o gpurun_out/r2_prof $CMD > gpurun_out/r2_ncu2.log 2>&1
CMDG="python bench.py --points 100e6 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-modes --ground grid"
$CMDG > gpurun_out/r2_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_grid_min|k_compact_xyz|k_minmax_f32' -s 3 -c 3 -o gpurun_out/r2_prof_grid $CMDG > gpurun_out/r2_ncu3.log 2>&1
tail -2 gpurun_out/r2_ncu2.log gpurun_out/r2_ncu3.log
ls -la gpurun_out/r2_prof*.ncu-rep gpurun_out/r2_launches.csv
CMDO="python tools/prof_obb.py 50e6"
$CMDO > gpurun_out/r2_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_obb' -c 1 -o gpurun_out/r2_prof_obb $CMDO > gpurun_out/r2_ncu4.log 2>&1
tail -2 gpurun_out/r2_ncu4.log
