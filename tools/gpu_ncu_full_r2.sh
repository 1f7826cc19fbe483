#!/bin/bash
# ncu --set full of the three kernels that lead the batch (rowstream<64>, halo pair<256>, rowstream2<16,HEAD>) inside the bench workload
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-library --no-kernel-table"
timeout 600 $CMD > gpurun_out/plain_full_r2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"conv_rowstream|conv_halo_pair_kernel<256" -s 300 -c 14 -o gpurun_out/prof_r2_top $CMD > gpurun_out/ncu_full_r2.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_full_r2.log
