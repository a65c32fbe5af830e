#!/bin/bash
# Round 2, GPU call 19 (two GPUs): the code as committed on two GPUs -- multi-GPU tests and the bench line.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multi_gpu.py -q -x > gpurun_out/r2zz_tests_n2.log 2>&1; echo "multi-gpu tests rc=$?"; tail -4 gpurun_out/r2zz_tests_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29573 \
   bench.py --gpus 2 > gpurun_out/r2zz_bench_n2.json 2> gpurun_out/r2zz_bench_n2.err; echo "bench n2 rc=$?"
grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r2zz_bench_n2.err | tail -3
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2zz_bench_n2.json').read().strip().splitlines()[-1])
    g = d['grad_eval']
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'error')}, 'e2e ms', d['e2e']['ms_per_step'])
    print({k: g.get(k) for k in ('seconds', 'pcg_iterations', 'ms_per_pcg_iteration', 'stage_s', 'u_rel_diff_vs_replicated_solve', 'halo_exchanges', 'error')})
    print((d.get('batch_eval') or {}).get('designs_per_s'), (g.get('u_err_estimate') or {}).get('value'))
except Exception as e:
    print('parse failed', e)
PY
