#!/bin/bash
# Round 2, GPU call 6 (eight GPUs): bench lines at N = 8 and N = 4 (distributed peer-memory solve incl. distributed
# numeric setup, config-5 batch leg), phase timing of the numeric setup, two multi-GPU parity tests at world 8.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
for N in 8 4; do
  JSSO_MG_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N \
     bench.py --gpus $N --steps 10 > gpurun_out/r2c6_bench_n$N.json 2> gpurun_out/r2c6_bench_n$N.err; echo "bench n$N rc=$?"
  grep JSSO_MG_TIMING gpurun_out/r2c6_bench_n$N.err | tail -3
  grep -v "JSSO_MG_TIMING\|OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r2c6_bench_n$N.err | tail -3
  python - $N <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/r2c6_bench_n%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
    g = d['grad_eval']
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'error')}, d['e2e']['ms_per_step'])
    print({k: g.get(k) for k in ('seconds', 'seconds_each', 'pcg_iterations', 'ms_per_pcg_iteration', 'stage_s', 'u_rel_diff_vs_replicated_solve', 'halo_exchanges', 'error')})
    print(g['multigrid']['distributed'])
    print(d.get('batch_eval'))
except Exception as e:
    print('parse failed', e)
PY
done
timeout 600 python -m pytest tests/test_multi_gpu.py -q -x -k "partitioned_equals_single and 8 or (distributed_multigrid and 8-500 and setup)" > gpurun_out/r2c6_tests.log 2>&1; echo "world-8 tests rc=$?"; tail -4 gpurun_out/r2c6_tests.log
