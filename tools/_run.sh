bash tools/ncu_launches.sh r01v7
NCU_COUNT=8 bash tools/ncu_full.sh gemm_v2 gemm_v2_v7 470
