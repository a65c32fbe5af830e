#!/bin/bash
# Round 2, GPU call 22 (one GPU): last check of the code as committed (assembly refactored into range helpers):
# full GPU test suite, smoke, short bench.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r2end_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2end_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --no-cpu-baseline --batch-designs 0 --topo-iters 0 > gpurun_out/r2end_bench_n1.json 2> gpurun_out/r2end_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2end_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2end_bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'kernel_ms', 'gpu_launches')}); print(d['e2e']['ms_per_step']); print(d['m2']['seconds_per_eval'], d['m2']['pcg_iterations'], d['m2']['e2e']['seconds'])
PY
