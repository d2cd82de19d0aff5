#!/bin/bash
# round-2 ncu captures (one gpurun call): dense kernels and HBM-bound kernels, raw metrics exported to CSV
mkdir -p gpurun_out
python tools/prof_kernels.py dense > gpurun_out/prof_dense_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_ln|gemm_v2|attn_bwd_tc|attn_tc_fwd|attn_dq_finish|layernorm_fwd" -c 40 -o gpurun_out/prof_r02_dense -f python tools/prof_kernels.py dense > gpurun_out/ncu_r02_dense.log 2>&1
echo "ncu dense exit $?"
python tools/prof_kernels.py hbm > gpurun_out/prof_hbm_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:"layernorm|kd_loss_fwd|kd_loss_bwd|bn_|tokenize|im2col|colreduce|token_mean|cast_|batch_sum|param_prep" -c 120 -o gpurun_out/prof_r02_hbm -f python tools/prof_kernels.py hbm > gpurun_out/ncu_r02_hbm.log 2>&1
echo "ncu hbm exit $?"
for t in dense hbm; do
  ncu -i gpurun_out/prof_r02_$t.ncu-rep --page raw --csv > gpurun_out/prof_r02_${t}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/prof_r02_dense.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_r02_dense_source.csv 2>/dev/null
gzip -f gpurun_out/prof_r02_dense_source.csv
ls -la gpurun_out/prof_r02*
for t in dense hbm; do
  sz=$(stat -c %s gpurun_out/prof_r02_$t.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 25000000 ]; then rm -f gpurun_out/prof_r02_$t.ncu-rep; echo "report $t dropped ($sz bytes)"; fi
done
