#!/bin/bash
# Round 2, GPU call 17 (one GPU): hierarchy stopping at <= 256 nodes with the blocked dense inverse of the coarsest
# level against the hierarchy down to <= 64 nodes: GPU tests, A/B, setup phases.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigrid.py tests/test_large_parity.py tests/test_api_gpu.py -q -m gpu -x > gpurun_out/r2aa_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2aa_tests.log
MG_SWITCH_DEGREES=1 MG_SWITCH_CONFIGS='[{}, {"MAX_COARSE": "64"}, {"MAX_COARSE": "1000"}]' \
  timeout 600 python scripts/mg_switches.py 1024 1e-8 > gpurun_out/r2aa_mg_switches.txt 2>&1; echo "switches rc=$?"; grep "MG_SWITCH\|symbolic" gpurun_out/r2aa_mg_switches.txt | cut -c1-330
JSSO_MG_TIMING=1 timeout 600 python scripts/mg_profile.py 1024 3 1 setup > gpurun_out/r2aa_setup_phases.txt 2>&1; echo "phases rc=$?"; grep JSSO_MG_TIMING gpurun_out/r2aa_setup_phases.txt | tail -2 | cut -c1-900
