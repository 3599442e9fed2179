#!/bin/bash
# developer tool (GPU box): guarded smoke -> GPU tests -> A/B sweep; every step under `timeout`,
# logs in gpurun_out/ so that a hang costs a bounded amount of box time
mkdir -p gpurun_out
for s in 64 32 128 256; do
  timeout 180 python tools/prof_stack.py $s 300 > gpurun_out/smoke_$s.log 2>&1; rc=$?
  echo "smoke $s rc=$rc $(tail -1 gpurun_out/smoke_$s.log | cut -c1-100)"
  if [ $rc -ne 0 ]; then echo "smoke failed - stopping"; exit 1; fi
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)"
