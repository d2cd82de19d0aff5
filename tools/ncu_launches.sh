#!/bin/bash
# launch list (device time per kernel) of a short bench run; $1 = output tag, rest = extra bench args
tag=$1; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e $*"
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-1500} -c ${NCU_COUNT:-1000} --csv \
    --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_$tag.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/plain_$tag.log | cut -c1-300; wc -l gpurun_out/launches_$tag.csv
