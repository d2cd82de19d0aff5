#!/bin/bash
# one `ncu --set full` capture of a kernel family inside a short bench run; $1 = kernel regex, $2 = tag, $3 = skip count
# The raw-metric and source pages are exported to CSV on the GPU box; the .ncu-rep itself is only kept when small
# (gpurun brings back at most 64 MiB).
mkdir -p gpurun_out
CMD="${NCU_CMD:-python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e --no-graph ${BENCH_ARGS:-}}"
$CMD > gpurun_out/plain_full_$2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-60} -c ${NCU_COUNT:-6} -o gpurun_out/prof_$2 -f $CMD > gpurun_out/ncu_full_$2.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_$2.ncu-rep --page raw --csv > gpurun_out/prof_$2_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$2.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_$2_source.csv 2>/dev/null
gzip -f gpurun_out/prof_$2_source.csv
ls -la gpurun_out/prof_$2*
sz=$(stat -c %s gpurun_out/prof_$2.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 30000000 ]; then rm -f gpurun_out/prof_$2.ncu-rep; echo "report dropped ($sz bytes)"; fi
