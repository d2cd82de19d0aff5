#!/bin/bash
# A/B matrix for the GEMM: cta_group {1,2} x tile width x epilogue skipped or not (B200_GEMM_DBG=1 drains TMEM only)
for two in 0 1; do for bn in 128 192 256; do for dbg in 0 1; do
  echo "=== 2CTA=$two BN=$bn DBG=$dbg"
  B200_GEMM_2CTA=$two B200_GEMM_BN=$bn B200_GEMM_DBG=$dbg timeout 300 python tools/gemm_bench.py ${GEMM_BENCH_ARGS:-vits} 2>&1 | tail -8
done; done; done
