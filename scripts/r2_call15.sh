#!/bin/bash
# Round 2, GPU call 15 (one GPU): numeric multigrid setup by phase (JSSO_MG_TIMING=1) with the six-threads-per-block
# Galerkin product kernel and the power iterations on the reduced-precision storage; multigrid GPU tests; short bench.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigrid.py tests/test_large_parity.py -q -m gpu -x > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2x_tests.log
JSSO_MG_TIMING=1 timeout 600 python scripts/mg_profile.py 1024 3 1 setup > gpurun_out/r2x_setup_phases.txt 2>&1; echo "phases rc=$?"; grep JSSO_MG_TIMING gpurun_out/r2x_setup_phases.txt | cut -c1-900
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r2x_launches_setup.csv python scripts/mg_profile.py 1024 3 1 setup > gpurun_out/r2x_mgprof.log 2>&1; echo "launch list rc=$?"
python scripts/launch_summary.py gpurun_out/r2x_launches_setup.csv > gpurun_out/r2x_launches_numeric_setup.txt; head -24 gpurun_out/r2x_launches_numeric_setup.txt
python bench.py --no-cpu-baseline --batch-designs 0 --topo-iters 0 > gpurun_out/r2x_bench_n1.json 2> gpurun_out/r2x_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2x_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2x_bench_n1.json').read().strip().splitlines()[-1])
print(d['m2']); print(d['grad_eval']['stage_s'], d['grad_eval']['seconds_each'])
PY
