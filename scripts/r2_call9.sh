#!/bin/bash
# Round 2, GPU call 9 (one GPU): record staging of assemble_tasks_kernel by bulk async copies (cp.async.bulk + mbarrier)
# against per-lane cp.async (A/B builds), parity tests of the assembly, the config-4 loop for 100 iterations.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "assembly or quad_ke or beam_ke or value_and_grad or full_size" > gpurun_out/r2c9_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c9_tests.log
for lib in libjsso.so libjsso_nobulk.so; do
  for n in 512 1024; do
    echo "$lib $n: $(JSSO_LIB=$PWD/jaxsso_b200/$lib timeout 300 python scripts/asm_time.py $n 2>&1 | tail -1)"
  done
done | tee gpurun_out/r2c9_asm_ab.txt
timeout 300 python scripts/topo_shape_512.py 512 100 2>&1 | tail -1 | tee gpurun_out/r2c9_topo100.json
