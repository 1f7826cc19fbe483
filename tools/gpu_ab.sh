#!/bin/bash
mkdir -p gpurun_out
for d in 1 2 4; do
echo "=== FLUSH=$d"; WSI_STREAM_FLUSH=$d WSI_CONV_TRACE=1 timeout 600 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/conv_trace_fl$d.log 2>&1; echo "exit $?"; grep -E "iter 2|64->64|32->32|16->16" gpurun_out/conv_trace_fl$d.log | cut -c1-120
done
