#!/bin/bash
# Round 2, GPU call 13 (eight GPUs): bench lines at N = 8 and 4 with the final code (ghost-row recomputation of the
# corrected iterate, power iterations on the reduced-precision storage).
set -u
mkdir -p gpurun_out
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$N \
     bench.py --gpus $N > gpurun_out/r2v_bench_n$N.json 2> gpurun_out/r2v_bench_n$N.err; echo "bench n$N rc=$?"
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r2v_bench_n$N.err | tail -2
  python - $N <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/r2v_bench_n%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
    g = d['grad_eval']
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'error')}, 'e2e ms', d['e2e']['ms_per_step'])
    print({k: g.get(k) for k in ('seconds', 'pcg_iterations', 'ms_per_pcg_iteration', 'stage_s', 'u_rel_diff_vs_replicated_solve', 'halo_exchanges', 'error')})
    print((d.get('batch_eval') or {}).get('designs_per_s'), (g.get('u_err_estimate') or {}).get('value'))
except Exception as e:
    print('parse failed', e)
PY
done
