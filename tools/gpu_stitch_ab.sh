#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 3 4 5; do
  echo "=== variant $v"; WSI_STITCH_VARIANT=$v timeout 300 python tools/perf_probe.py 12000 512 128 unet 2>&1 | grep -E "iter 2|stitch|gather" | tail -3
done
