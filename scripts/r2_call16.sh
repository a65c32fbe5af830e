#!/bin/bash
# Round 2, GPU call 16 (one GPU): the smallest multigrid levels as one thread-block-cluster kernel (since removed)
# (mg_tail_cluster_kernel) -- GPU tests, A/B against separate graph nodes over cluster size and row threshold, launch list.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigrid.py tests/test_large_parity.py -q -m gpu -x > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2y_tests.log
MG_SWITCH_DEGREES=1 MG_SWITCH_CONFIGS='[{}, {"JSSO_MG_CLUSTER": "0"}, {"JSSO_MG_CLUSTER_SIZE": "8"}, {"JSSO_MG_CLUSTER_ROWS": "20000"}, {"JSSO_MG_CLUSTER_ROWS": "500"}, {"JSSO_MG_CLUSTER_SIZE": "4"}]' \
  timeout 600 python scripts/mg_switches.py 1024 1e-8 > gpurun_out/r2y_mg_switches.txt 2>&1; echo "switches rc=$?"; grep MG_SWITCH gpurun_out/r2y_mg_switches.txt | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r2y_launches_mg.csv python scripts/mg_profile.py 1024 3 > gpurun_out/r2y_mgprof.log 2>&1; echo "launch list rc=$?"
python scripts/launch_sequence.py gpurun_out/r2y_launches_mg.csv 4 30 > gpurun_out/r2y_launches_mg_iteration.txt; tail -32 gpurun_out/r2y_launches_mg_iteration.txt
