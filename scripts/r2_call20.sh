#!/bin/bash
# Round 2, GPU call 20 (one GPU): `ncu --set full` of the level-0 kernels of the numeric multigrid setup (Galerkin
# products, prolongator smoothing, block-Jacobi scaling of the matrix) -- what bounds the 21 ms that are left.
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:'mg_block_product_kernel|mg_smooth_prolongator_kernel|scale_blocks_kernel|mg_scale_cols_kernel|mg_transpose_blocks_kernel' -c 5 \
   -o gpurun_out/r2ab_setup -f python scripts/mg_profile.py 1024 3 1 setup > gpurun_out/r2ab_mgprof_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/r2ab_mgprof_full.log
ncu -i gpurun_out/r2ab_setup.ncu-rep --page raw --csv > gpurun_out/r2ab_setup_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2ab_setup_raw.csv > gpurun_out/r2ab_ncu_numeric_setup_level0.txt 2>&1
rm -f gpurun_out/r2ab_setup.ncu-rep
grep -E "=====|time_duration|dram__bytes|lts__t_bytes|warps_active|long_scoreboard|lg_throttle|mio_throttle|issue_active|dram_throughput" gpurun_out/r2ab_ncu_numeric_setup_level0.txt | head -70
