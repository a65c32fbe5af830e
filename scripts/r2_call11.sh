#!/bin/bash
# Round 2, GPU call 11 (one GPU): the coarse tail as one cooperative kernel (mg_tail_kernel, since removed) on hardware -- full GPU
# test suite, A/B against the graph of kernels and over the grid size, launch list of one iteration, the bench line.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2t_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2t_tests.log
MG_SWITCH_DEGREES=1 MG_SWITCH_CONFIGS='[{}, {"JSSO_MG_TAIL": "0"}, {"JSSO_MG_TAIL_BLOCKS": "148"}, {"JSSO_MG_TAIL_BLOCKS": "296"}, {"JSSO_MG_TAIL_BLOCKS": "444"}]' \
  timeout 600 python scripts/mg_switches.py 1024 1e-8 > gpurun_out/r2t_mg_switches.txt 2>&1; echo "switches rc=$?"; grep MG_SWITCH gpurun_out/r2t_mg_switches.txt | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r2t_launches_mg.csv python scripts/mg_profile.py 1024 3 > gpurun_out/r2t_mgprof.log 2>&1; echo "launch list rc=$?"
python scripts/launch_sequence.py gpurun_out/r2t_launches_mg.csv 4 16 > gpurun_out/r2t_launches_mg_iteration.txt; tail -18 gpurun_out/r2t_launches_mg_iteration.txt
python bench.py > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2t_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2t_bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'kernel_ms', 'gpu_launches', 'clocks')})
print(d['e2e']); print(d['m2']); print(d['roofline_pcg_iteration'])
print(d['grad_eval']['stage_s'], d['grad_eval']['seconds_each'], d['grad_eval'].get('u_err_estimate')); print(d.get('batch_eval')); print(d.get('topo_eval'))
PY
