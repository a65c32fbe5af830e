#!/bin/bash
# First GPU call of round 2 (needs >= 2 GPUs for the distributed part; on 1 GPU those tests skip):
#   gpurun --gpus 2 --timeout 1500 -- bash scripts/round2_first_call.sh
# 1. the verified GPU tests (must stay green), 2. the tests of code written after round 1's GPU budget
# (row-range distributed multigrid, binary16 fine level), 3. bench lines with / without the new switches.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r2_tests_verified.log 2>&1
echo "verified tests rc=$?"
JSSO_RUN_UNVERIFIED=1 timeout 1200 python -m pytest tests -q -m gpu -k "fp16 or distributed_multigrid" > gpurun_out/r2_tests_unverified.log 2>&1
echo "unverified tests rc=$?"
tail -5 gpurun_out/r2_tests_verified.log gpurun_out/r2_tests_unverified.log
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err
JSSO_MG_FP16=1 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r2_b1_fp16.json 2> gpurun_out/r2_b1_fp16.err
JSSO_MG_ASYNC=4 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/r2_b1_async.json 2> gpurun_out/r2_b1_async.err
python -c "import json; d=json.loads(open('gpurun_out/r2_b1_async.json').read().strip().splitlines()[-1]); print('async', d['grad_eval'])"
JSSO_MG_GRAPH=1 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/r2_b1_graph.json 2> gpurun_out/r2_b1_graph.err
python -c "import json; d=json.loads(open('gpurun_out/r2_b1_graph.json').read().strip().splitlines()[-1]); print('graph', d['grad_eval'])"
JSSO_MG_GRAPH=1 JSSO_MG_ASYNC=4 JSSO_MG_FP16=1 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/r2_b1_all.json 2> gpurun_out/r2_b1_all.err
python -c "import json; d=json.loads(open('gpurun_out/r2_b1_all.json').read().strip().splitlines()[-1]); print('graph+async+fp16', d['grad_eval'])"
for K in 2 4 8; do   # chunked host pipeline of the e2e leg (opt-in): compare e2e.ms_per_step
  JSSO_E2E_CHUNKS=$K python bench.py --steps 10 --no-cpu-baseline --no-solve > gpurun_out/r2_b1_e2e$K.json 2> gpurun_out/r2_b1_e2e$K.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/r2_b1_e2e$K.json').read().strip().splitlines()[-1]); print('e2e chunks $K', d['e2e'])"
done
python - <<'PY'
import json
for f in ('gpurun_out/r2_b1.json', 'gpurun_out/r2_b1_fp16.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['grad_eval'])
    except Exception as e:
        print(f, 'failed', e)
PY
# peer-memory variant of the distributed multigrid exchanges + the partitioned checks (2 GPUs)
NG=$(python -c "from jaxsso_b200 import _native as n; print(n.lib().jsso_device_count())")
if [ "$NG" -ge 2 ]; then
  for extra in "" "p2p"; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
      scripts/dist_mg_check.py 96 500 1 $extra 2>&1 | grep DIST_MG_CHECK
  done
fi
if [ "$NG" -ge 2 ]; then
  for extra in "" "--dist-mg"; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
      bench.py --gpus 2 --steps 10 $extra > "gpurun_out/r2_b2${extra}.json" 2> "gpurun_out/r2_b2${extra}.err"
    tail -c 1500 "gpurun_out/r2_b2${extra}.json"
  done
fi
