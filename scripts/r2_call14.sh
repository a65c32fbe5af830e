#!/bin/bash
# Round 2, GPU call 14 (one GPU): full GPU test suite and the bench line with the code as committed (graph for the
# coarse levels, power iterations on the reduced-precision storage, host-buffer entry point with direct DMA / threaded
# staging).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2w_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2w_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2w_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2w_bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'kernel_ms', 'gpu_launches', 'clocks')})
print(d['e2e']); print(d['m2']); print(d['roofline_pcg_iteration']['frac'], d['roofline']['frac'], d['roofline_assembly']['frac'], d['roofline_adjoint']['frac'], d['roofline_spmv']['frac'])
print(d['grad_eval']['stage_s'], d['grad_eval']['seconds_each']); print(d.get('batch_eval', {}).get('designs_per_s')); print(d.get('topo_eval', {}).get('s_per_iteration')); print(d.get('cpu_baseline'))
PY
