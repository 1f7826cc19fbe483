#!/bin/bash
mkdir -p gpurun_out
for v in "WSI_NONE=1" "WSI_NO_HALO_UP2=1" "WSI_NONE=1" "WSI_NO_HALO_UP2=1"; do
  echo "=== [$v]"
  env $v WSI_CONV_TRACE=1 timeout 300 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/conv_trace_ab.log 2>&1; echo "exit $?"
  grep -E "iter 2|up2 BN" gpurun_out/conv_trace_ab.log | cut -c1-100
done
