#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/perf_probe.py 4096 512 128 unet"
timeout 300 $CMD > gpurun_out/plain_stitch.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stitch_finalise_seg|gather_kernel" -s 6 -c 6 -o gpurun_out/prof_stitch $CMD > gpurun_out/ncu_stitch.log 2>&1
echo "ncu exit $?"; tail -n 5 gpurun_out/ncu_stitch.log; tail -n 12 gpurun_out/plain_stitch.log
