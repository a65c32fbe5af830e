#!/bin/bash
# one GPU call: parity of the task assembly path, timing of both paths
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --no-solve --no-cpu-baseline > gpurun_out/b_tasks.json 2>gpurun_out/b_tasks.err
tail -c 400 gpurun_out/b_tasks.err
python - <<'PY'
import json
for f in ("b_tasks",):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1])
        print(f, d["kernel_ms"], d["ms_per_step"], d["e2e"]["ms_per_step"])
    except Exception as e:
        print(f, 'failed', e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_asm.csv python scripts/prof_asm.py 1024 2>&1 | tail -1
python scripts/launch_summary.py gpurun_out/launches_asm.csv
