#!/bin/bash
# The single-GPU `-m gpu` suite against the EMULATED library (tests/emu): a regression check of the host driver and
# the kernels' logic for a container without a GPU.  ~25 minutes on 8 cores; not part of the CPU suite.
#   bash scripts/gpu_tests_on_emulator.sh [extra pytest args]
set -e
cd "$(dirname "$0")/.."
LIB=$(python -c "import sys; sys.path.insert(0, 'tests/emu'); import build_emu; print(build_emu.build_api())")
JSSO_LIB=$LIB JSSO_RUN_UNVERIFIED=1 JSSO_FULL_SIZE=10 python -m pytest tests -m gpu -q --timeout 900 \
  -k "not multi_gpu and not partitioned and not distributed_multigrid" "$@"
