#!/bin/bash
# scratch A/B: row-stream dependent-MMA-chain experiment (WSI_STREAM_DBG=4; results are garbage)
mkdir -p gpurun_out
for d in 0 4; do
  echo "=== WSI_STREAM_DBG=$d"
  WSI_STREAM_DBG=$d WSI_CONV_TRACE=1 timeout 300 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/conv_trace_stdbg$d.log 2>&1; echo "exit $?"
  grep -E "iter 2|16->16" gpurun_out/conv_trace_stdbg$d.log | cut -c1-110
done
