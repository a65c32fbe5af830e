#!/bin/bash
# Round 2, GPU call 7 (eight GPUs): N = 8 bench line after the vectorised one-block exchange kernel.
set -u
mkdir -p gpurun_out
JSSO_MG_TIMING=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29578 \
   bench.py --gpus 8 --steps 10 --batch-designs 0 > gpurun_out/r2c8_bench_n8.json 2> gpurun_out/r2c8_bench_n8.err; echo "bench n8 rc=$?"
grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r2c8_bench_n8.err | tail -3
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2c8_bench_n8.json').read().strip().splitlines()[-1])
    g = d['grad_eval']
    print({k: g.get(k) for k in ('seconds', 'seconds_each', 'pcg_iterations', 'ms_per_pcg_iteration', 'stage_s', 'u_rel_diff_vs_replicated_solve', 'error')})
except Exception as e:
    print('parse failed', e)
PY
