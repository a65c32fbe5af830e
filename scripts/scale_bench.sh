#!/bin/bash
# bench.py at N = 2, 4, 8 GPUs of one box (the driver's own scaling run does the same): gpurun --gpus 8 -- bash scripts/scale_bench.sh
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540 + n)) \
      bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r1d_bench_n$n.json 2> gpurun_out/r1d_bench_n$n.err
  tail -c 200 gpurun_out/r1d_bench_n$n.err | grep -v OMP_NUM
  python - <<PY
import json
d = json.loads(open('gpurun_out/r1d_bench_n$n.json').read().strip().splitlines()[-1])
print($n, {k: d[k] for k in ('value', 'ms_per_step', 'kernel_ms')}, d['grad_eval']['seconds'], d['grad_eval']['solve'], d['e2e']['ms_per_step'])
PY
done
