#!/bin/bash
# Round 2, GPU call 1 (one GPU): baseline tests, the unmeasured opt-in switches of round 1, rtol sweep at the
# benchmark's own size, FP64 peak, ncu of the multigrid-PCG iteration.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2c1_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c1_tests.log
python -c "
from jaxsso_b200 import _native as nat
print('FP64_PEAK', nat.fp64_peak(0, 3.0))" 2>&1 | tee gpurun_out/r2c1_fp64.txt
timeout 600 python scripts/mg_switches.py 1024 1e-8 > gpurun_out/r2c1_switches.log 2>&1; echo "switches rc=$?"; grep MG_SWITCH gpurun_out/r2c1_switches.log
timeout 600 python scripts/rtol_sweep.py 1024 1 > gpurun_out/r2c1_rtol.log 2>&1; echo "rtol rc=$?"; grep -v RTOL_SWEEP gpurun_out/r2c1_rtol.log | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r2c1_launches_mg.csv python scripts/mg_profile.py 1024 3 > gpurun_out/r2c1_mgprof.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2c1_mg -f \
   python scripts/mg_profile.py 1024 2 > gpurun_out/r2c1_mgprof_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r2c1_mg.ncu-rep --page raw --csv > gpurun_out/r2c1_mg_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2c1_mg_raw.csv > gpurun_out/r2c1_ncu_mg_1024.txt 2>&1
rm -f gpurun_out/r2c1_mg.ncu-rep   # the raw csv + summary come back; the report itself is large
JSSO_E2E_CHUNKS=4 python bench.py --steps 10 --no-cpu-baseline --no-solve > gpurun_out/r2c1_b_e2e4.json 2> gpurun_out/r2c1_b_e2e4.err
python -c "
import json
d = json.loads(open('gpurun_out/r2c1_b_e2e4.json').read().strip().splitlines()[-1]); print('e2e chunks 4', d['e2e'], d['ms_per_step'])"
python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r2c1_bench.json
