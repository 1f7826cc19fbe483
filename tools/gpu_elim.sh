#!/bin/bash
# elimination experiment on the row kernels (debug-switch build ab/libG.so; results are garbage, times are what is left)
export WSI_B200_LIB=$PWD/ab/libG.so
for d in 0 1 2 3; do
  echo "=== WSI_STREAM_DBG=$d WSI_UP_DBG=$([ $d == 1 ] && echo 1 || echo 0)"
  WSI_STREAM_DBG=$d WSI_UP_DBG=$([ $d == 1 ] && echo 1 || echo 0) timeout 200 python tools/op_probe.py 8192 "rowstream,upstream" 2>&1 | tail -12
done
