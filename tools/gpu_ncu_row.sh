#!/bin/bash
# ncu --set full of the high-resolution decoder kernels (two-lane row-stream, x2 row-stream)
mkdir -p gpurun_out
CMD="python tools/perf_probe.py 4096 512 128 unet"
timeout 600 $CMD > gpurun_out/plain_row.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_rowstream2|conv_upstream" -s 8 -c 4 -o gpurun_out/prof_row $CMD > gpurun_out/ncu_row.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_row.log
