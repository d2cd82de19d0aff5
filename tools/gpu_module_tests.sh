#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary_mod.txt
for k in "$@"; do
  timeout 900 python -m pytest tests/test_gpu_modules.py -q --tb=short -k "$k" > "gpurun_out/mod_$k.log" 2>&1
  echo "$k exit $?" >> gpurun_out/summary_mod.txt
  tail -n 25 "gpurun_out/mod_$k.log" >> gpurun_out/summary_mod.txt
done
cat gpurun_out/summary_mod.txt
