#!/bin/bash
# scratch: same-box A/B of two builds of the library (WSI_B200_LIB selects the .so)
mkdir -p gpurun_out
for lib in "" "$PWD/gpurun_out_old_lib.so" "" "$PWD/gpurun_out_old_lib.so"; do
  echo "=== lib [$lib]"
  WSI_B200_LIB=$lib timeout 300 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/probe_ab.log 2>&1; echo "exit $?"
  grep -E "iter [12]|classes_hist" gpurun_out/probe_ab.log | cut -c1-100
  grep -E "iter 2" -A5 gpurun_out/probe_ab.log | grep "conv  "
done
