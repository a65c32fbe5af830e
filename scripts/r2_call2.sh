#!/bin/bash
# Round 2, GPU call 2 (one GPU): the fused multigrid-PCG iteration -- parity tests, A/B of its switches, launch list
# and ncu of the iteration, bench line.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2c2_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2c2_tests.log
timeout 600 python scripts/mg_switches.py 1024 1e-8 > gpurun_out/r2c2_switches.log 2>&1; echo "switches rc=$?"; grep MG_SWITCH gpurun_out/r2c2_switches.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r2c2_launches_mg.csv python scripts/mg_profile.py 1024 3 > gpurun_out/r2c2_mgprof.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2c2_mg -f \
   python scripts/mg_profile.py 1024 2 > gpurun_out/r2c2_mgprof_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r2c2_mg.ncu-rep --page raw --csv > gpurun_out/r2c2_mg_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2c2_mg_raw.csv > gpurun_out/r2c2_ncu_mg_1024.txt 2>&1
rm -f gpurun_out/r2c2_mg.ncu-rep
python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err; echo "bench rc=$?"
python -c "
import json
d = json.loads(open('gpurun_out/r2c2_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'kernel_ms')}); print(d['e2e']); print(d['grad_eval'])"
