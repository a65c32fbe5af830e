#!/bin/bash
# Round 2, GPU call 23 (two GPUs): last check of the distributed path with the code as committed.
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29575 \
   bench.py --gpus 2 --batch-designs 0 > gpurun_out/r2end_bench_n2.json 2> gpurun_out/r2end_bench_n2.err; echo "bench n2 rc=$?"
grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r2end_bench_n2.err | tail -2
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2end_bench_n2.json').read().strip().splitlines()[-1])
g = d['grad_eval']
print({k: d.get(k) for k in ('value', 'ms_per_step', 'error')})
print({k: g.get(k) for k in ('seconds', 'pcg_iterations', 'ms_per_pcg_iteration', 'stage_s', 'u_rel_diff_vs_replicated_solve', 'error')})
PY
timeout 300 python -m pytest tests/test_multi_gpu.py -q -x -k "setup" > gpurun_out/r2end_tests_n2.log 2>&1; echo "multi-gpu setup tests rc=$?"; tail -2 gpurun_out/r2end_tests_n2.log
