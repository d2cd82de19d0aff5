#!/bin/bash
# Runs each kernel test group in its own process (a faulting kernel poisons only its own CUDA context).
mkdir -p gpurun_out
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for k in "$@"; do
  timeout 600 python -m pytest tests/test_gpu_kernels.py -q --tb=short -k "$k" > "gpurun_out/$k.log" 2>&1
  echo "$k exit $?" >> gpurun_out/summary.txt
  tail -n 3 "gpurun_out/$k.log" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
