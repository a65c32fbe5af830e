#!/bin/bash
# Round 2, GPU call 5 (two GPUs): distributed numeric setup + fused all-reduces + one-kernel exchanges on hardware.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multi_gpu.py -q -x > gpurun_out/r2c5_tests.log 2>&1; echo "multi-gpu tests rc=$?"; tail -6 gpurun_out/r2c5_tests.log
for extra in "" "--replicated-setup"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
     bench.py --gpus 2 --steps 10 --batch-designs 0 $extra > "gpurun_out/r2c5_bench_n2$extra.json" 2> "gpurun_out/r2c5_bench_n2$extra.err"; echo "bench n2 $extra rc=$?"; tail -3 "gpurun_out/r2c5_bench_n2$extra.err"
  python - "$extra" <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/r2c5_bench_n2%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
    g = d['grad_eval']
    print({k: g.get(k) for k in ('seconds', 'pcg_iterations', 'ms_per_pcg_iteration', 'stage_s', 'u_rel_diff_vs_replicated_solve', 'error')})
except Exception as e:
    print('parse failed', e)
PY
done
