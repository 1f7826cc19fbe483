#!/bin/bash
# ncu --set full of the memory-bound side kernels (gather, stitch, max-pool)
mkdir -p gpurun_out
CMD="python tools/perf_probe.py 4096 512 128 unet"
timeout 600 $CMD > gpurun_out/plain_misc.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gather_kernel|stitch_seg_kernel|maxpool_planar" -s 6 -c 6 -o gpurun_out/prof_misc $CMD > gpurun_out/ncu_misc.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_misc.log
