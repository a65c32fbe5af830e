#!/bin/bash
# Round 2, GPU call 3 (one GPU): row-pair SpMV for the V-cycle products -- solver tests, A/B, launch lists (iteration and
# setup), ncu of the iteration, the new bench line.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "multigrid or large_parity or gpu_parity or api_gpu" > gpurun_out/r2c3_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2c3_tests.log
timeout 600 python scripts/mg_switches.py 1024 1e-8 > gpurun_out/r2c3_switches.log 2>&1; echo "switches rc=$?"; grep MG_SWITCH gpurun_out/r2c3_switches.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r2c3_launches_mg.csv python scripts/mg_profile.py 1024 3 > gpurun_out/r2c3_mgprof.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r2c3_launches_setup.csv python scripts/mg_profile.py 1024 1 1 setup > gpurun_out/r2c3_setupprof.log 2>&1; echo "setup launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2c3_mg -f \
   python scripts/mg_profile.py 1024 2 > gpurun_out/r2c3_mgprof_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r2c3_mg.ncu-rep --page raw --csv > gpurun_out/r2c3_mg_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2c3_mg_raw.csv > gpurun_out/r2c3_ncu_mg_1024.txt 2>&1
rm -f gpurun_out/r2c3_mg.ncu-rep
python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c3_bench.err
python -c "
import json
d = json.loads(open('gpurun_out/r2c3_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'kernel_ms')}); print(d['e2e']); print(d['grad_eval']); print(d.get('m2')); print(d.get('roofline_pcg_iteration')); print(d['roofline_adjoint'])"
