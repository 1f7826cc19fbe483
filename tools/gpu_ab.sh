#!/bin/bash
# scratch A/B: TMA A-tile row cost, strided NHWC rows vs contiguous chunk-planar rows (WSI_IGEMM_ALT; garbage results)
mkdir -p gpurun_out
for a in 0 1; do
  for d in 0 1; do
    echo "=== ALT=$a DBG=$d"
    if [ $a == 1 ]; then export WSI_IGEMM_ALT=1; else unset WSI_IGEMM_ALT; fi
    WSI_IGEMM_DBG=$d WSI_CONV_TRACE=1 timeout 300 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/conv_trace_alt$a$d.log 2>&1; echo "exit $?"
    grep -E "iter 2|128->128  @64x64|256->256  @32x32 BN256|512->512  @16x16 BN256 BK64 x2   " gpurun_out/conv_trace_alt$a$d.log | cut -c1-100
  done
done
