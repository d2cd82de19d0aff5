#!/bin/bash
# one `ncu --set full` capture of a kernel family inside a short bench run; $1 = kernel regex, $2 = tag, $3 = skip count
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e --no-graph ${BENCH_ARGS:-}"
$CMD > gpurun_out/plain_full_$2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-60} -c ${NCU_COUNT:-6} -o gpurun_out/prof_$2 -f $CMD > gpurun_out/ncu_full_$2.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_$2.ncu-rep 2>/dev/null
