#!/bin/bash
mkdir -p gpurun_out
for d in 2 4 8 16; do
echo "=== conv trace STAGES=$d"; WSI_ROW_STAGES=$d WSI_CONV_TRACE=1 timeout 600 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/conv_trace_st$d.log 2>&1; echo "exit $?"; grep -E "iter 2|@256x256|@512x512" gpurun_out/conv_trace_st$d.log
done
