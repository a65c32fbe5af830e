#!/bin/bash
# Round 2, GPU call 21 (one GPU): host pipeline of jsso_assemble_adjoint_host with the assembly interleaved by task ranges
# between the adjoint's quad ranges (the e2e leg of bench.py), against the whole assembly first.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "host_buffer or pinned or assembly or value_and_grad" > gpurun_out/r2ac_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2ac_tests.log
for il in 1 0 1 0; do
  JSSO_E2E_INTERLEAVE=$il python bench.py --no-solve --no-cpu-baseline --batch-designs 0 --topo-iters 0 > gpurun_out/r2ac_bench_il$il.json 2> gpurun_out/r2ac_bench_il$il.err; echo "bench interleave=$il rc=$?"
  python - $il <<'PY'
import json, sys
d = json.loads(open('gpurun_out/r2ac_bench_il%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print('interleave', sys.argv[1], 'value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['pcie']['copies_serial_ms'], d['e2e']['pcie']['copy_bound_ms'])
PY
done
for K in 8 16; do
  JSSO_E2E_CHUNKS=$K python bench.py --no-solve --no-cpu-baseline --batch-designs 0 --topo-iters 0 > gpurun_out/r2ac_bench_k$K.json 2> gpurun_out/r2ac_bench_k$K.err
  python - $K <<'PY'
import json, sys
d = json.loads(open('gpurun_out/r2ac_bench_k%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print('chunks', sys.argv[1], 'e2e ms', d['e2e']['ms_per_step'])
PY
done
