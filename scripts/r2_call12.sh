#!/bin/bash
# Round 2, GPU call 12 (two GPUs): ghost-row recomputation of the corrected iterate (two exchanges fewer per
# iteration), the exchange kernel's fence on the flag threads only and the cooperative coarse-tail kernel as the
# replicated tail: tail-kernel A/B on one GPU, multi-GPU tests, bench at N = 2 (tail kernel / graph).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigrid.py -q -m gpu -x -k coarse_tail > gpurun_out/r2u_tests1.log 2>&1; echo "tail test rc=$?"; tail -3 gpurun_out/r2u_tests1.log
MG_SWITCH_DEGREES=1 MG_SWITCH_CONFIGS='[{}, {"JSSO_MG_TAIL": "0"}, {"JSSO_MG_TAIL_ROWS": "200000"}, {"JSSO_MG_TAIL_ROWS": "2000"}]' \
  timeout 600 python scripts/mg_switches.py 1024 1e-8 > gpurun_out/r2u_mg_switches.txt 2>&1; echo "switches rc=$?"; grep MG_SWITCH gpurun_out/r2u_mg_switches.txt | cut -c1-330
timeout 1500 python -m pytest tests/test_multi_gpu.py -q -x > gpurun_out/r2u_tests.log 2>&1; echo "multi-gpu tests rc=$?"; tail -6 gpurun_out/r2u_tests.log
for tail in 1 0; do
  extra=""; [ $tail = 0 ] && extra="--no-cpu-baseline --batch-designs 0 --topo-iters 0 --steps 5"
  JSSO_MG_TAIL=$tail timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$tail \
     bench.py --gpus 2 $extra > gpurun_out/r2u_bench_n2_tail$tail.json 2> gpurun_out/r2u_bench_n2_tail$tail.err; echo "bench n2 tail=$tail rc=$?"
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r2u_bench_n2_tail$tail.err | tail -3
  python - $tail <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/r2u_bench_n2_tail%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
    g = d['grad_eval']
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'error')}, 'e2e ms', d['e2e']['ms_per_step'])
    print({k: g.get(k) for k in ('seconds', 'pcg_iterations', 'ms_per_pcg_iteration', 'stage_s', 'u_rel_diff_vs_replicated_solve', 'halo_exchanges', 'error')})
    print((d.get('batch_eval') or {}).get('designs_per_s'), (g.get('u_err_estimate') or {}).get('value'))
except Exception as e:
    print('parse failed', e)
PY
done
