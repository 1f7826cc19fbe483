#!/bin/bash
# scratch A/B on one box: halo kernel for stride-2 / x2 convs on vs off
mkdir -p gpurun_out
for v in "" "WSI_NO_HALO_UP2=1" "WSI_NO_HALO_S2=1" "WSI_NO_HALO_UP2=1 WSI_NO_HALO_S2=1"; do
  echo "=== [$v]"
  env $v WSI_CONV_TRACE=1 timeout 300 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/conv_trace_ab.log 2>&1; echo "exit $?"
  grep -E "iter 2|conv3x3/s2|up2 BN(64|128|256)|128->128  @64x64 BN128" gpurun_out/conv_trace_ab.log | cut -c1-100
  grep -E "iter 2" -A5 gpurun_out/conv_trace_ab.log | grep "conv  " | tail -1
done
