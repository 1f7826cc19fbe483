#!/bin/bash
# scratch A/B: cost of tcgen05.commit in the two-lane row-stream kernel (WSI_STREAM_DBG=6 adds two per row; =1 no MMAs)
mkdir -p gpurun_out
for d in 0 6 1; do
  echo "=== WSI_STREAM_DBG=$d"
  WSI_STREAM_DBG=$d WSI_CONV_TRACE=1 timeout 300 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/conv_trace_stdbg$d.log 2>&1; echo "exit $?"
  grep -E "iter 2|16->16|32->32" gpurun_out/conv_trace_stdbg$d.log | cut -c1-110
done
echo "=== single lane"
WSI_STREAM_LANES1=1 WSI_CONV_TRACE=1 timeout 300 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/conv_trace_l1.log 2>&1; echo "exit $?"
grep -E "iter 2|16->16|32->32" gpurun_out/conv_trace_l1.log | cut -c1-110
