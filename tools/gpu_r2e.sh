#!/bin/bash
mkdir -p gpurun_out
for v in tma direct; do
  echo "=== variant $v"
  if [ $v == direct ]; then export WSI_STITCH_DIRECT=1; fi
  timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider -k "stitch or canvas or unaligned or golden or invariance or full_size or empty or stride" > gpurun_out/pytest_r2s_$v.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/pytest_r2s_$v.log | cut -c1-300
  timeout 300 python tools/perf_probe.py 12000 512 128 unet 2>&1 | grep -E "iter 2|stitch" | tail -2
done
