#!/bin/bash
# Round 2, GPU call 4 (two GPUs): multi-GPU parity tests (partitioned path, row-range distributed multigrid over NCCL and
# over peer memory), bench line at N = 2 (distributed peer-memory solve as the gradient evaluation, checked against
# a replicated solve; config-5 batch leg), and the N = 1 line with the re-tuned setup for comparison.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 1500 python -m pytest tests/test_multi_gpu.py -q -x > gpurun_out/r2c4_tests.log 2>&1; echo "multi-gpu tests rc=$?"; tail -6 gpurun_out/r2c4_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
   bench.py --gpus 2 --steps 10 > gpurun_out/r2c4_bench_n2.json 2> gpurun_out/r2c4_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r2c4_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2c4_bench_n2.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'error')}); print(d['e2e']); print(d['grad_eval']); print(d.get('batch_eval'))
except Exception as e:
    print('n2 parse failed', e)
PY
timeout 900 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r2c4_bench_n1.json 2> gpurun_out/r2c4_bench_n1.err; echo "bench n1 rc=$?"; tail -3 gpurun_out/r2c4_bench_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2c4_bench_n1.json').read().strip().splitlines()[-1])
    print(d['grad_eval']); print(d.get('batch_eval')); print(d.get('topo_eval'))
except Exception as e:
    print('n1 parse failed', e)
PY
