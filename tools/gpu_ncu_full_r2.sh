#!/bin/bash
# ncu --set full of selected kernels inside the bench workload.  usage: gpu_ncu_full_r2.sh <regex> <out-name> [skip] [count]
mkdir -p gpurun_out
RX=${1:-"conv_rowstream|conv_halo_pair_kernel<256"}; OUT=${2:-prof_r2_top}; SKIP=${3:-300}; CNT=${4:-14}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-library --no-kernel-table"
timeout 600 $CMD > gpurun_out/plain_full_r2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT -o gpurun_out/$OUT $CMD > gpurun_out/ncu_full_r2.log 2>&1
echo "ncu exit $?"; tail -n 2 gpurun_out/ncu_full_r2.log
