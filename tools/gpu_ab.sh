#!/bin/bash
mkdir -p gpurun_out
for d in 5; do
echo "=== conv trace DBG=$d"; WSI_ROW_PREFETCH=0 WSI_ROW_DBG=$d WSI_CONV_TRACE=1 timeout 600 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/conv_trace_dbg$d.log 2>&1; echo "exit $?"; grep -E "iter 2|@256x256|@512x512" gpurun_out/conv_trace_dbg$d.log; tail -3 gpurun_out/conv_trace_dbg$d.log
done
