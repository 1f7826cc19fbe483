#!/bin/bash
# scratch: same-box A/B of two builds of the library (WSI_B200_LIB selects the .so)
mkdir -p gpurun_out
for lib in "" "$PWD/gpurun_out_old_lib.so" "" "$PWD/gpurun_out_old_lib.so"; do
  echo "=== lib [$lib]"
  WSI_B200_LIB=$lib WSI_CONV_TRACE=1 timeout 300 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/conv_trace_ab.log 2>&1; echo "exit $?"
  grep -E "iter 2|128->128  @64x64|up2 BN(64|128) " gpurun_out/conv_trace_ab.log | cut -c1-100
done
