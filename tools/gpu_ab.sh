#!/bin/bash
# scratch A/B: do the high-resolution layers get faster when a batch's tensors fit in L2?  (per-layer ms / tiles)
mkdir -p gpurun_out
for b in 74 37 18 12 8; do
  echo "=== batch $b"
  WSI_CONV_TRACE=1 timeout 300 python tools/perf_probe.py 4096 512 128 unet $b > gpurun_out/conv_trace_b$b.log 2>&1; echo "exit $?"
  grep -E "iter 2|stem 7x7|BK16|192->64" gpurun_out/conv_trace_b$b.log | cut -c1-100
  grep -E "iter 2" -A4 gpurun_out/conv_trace_b$b.log | grep -E "maxpool|gather" | tail -2
done
