#!/bin/bash
# Round 2, GPU call 5 (two GPUs): distributed numeric setup + fused all-reduces + one-kernel exchanges + scaled coarse
# levels on hardware: solver / parity tests on one GPU, multi-GPU tests, bench at N = 2 (distributed vs replicated setup)
# and N = 1.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "multigrid or large_parity or forward_displacements or value_and_grad" > gpurun_out/r2c5_tests1.log 2>&1; echo "1-GPU tests rc=$?"; tail -4 gpurun_out/r2c5_tests1.log
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
timeout 900 python bench.py --steps 10 --no-cpu-baseline --batch-designs 0 --topo-iters 0 > gpurun_out/r2c5_bench_n1.json 2> gpurun_out/r2c5_bench_n1.err; echo "bench n1 rc=$?"; tail -3 gpurun_out/r2c5_bench_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2c5_bench_n1.json').read().strip().splitlines()[-1])
    g = d['grad_eval']
    print({k: g.get(k) for k in ('seconds', 'pcg_iterations', 'ms_per_pcg_iteration', 'stage_s', 'u_err_estimate', 'error')})
    print(d.get('roofline_pcg_iteration'))
except Exception as e:
    print('n1 parse failed', e)
PY
