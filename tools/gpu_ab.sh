#!/bin/bash
# scratch A/B
mkdir -p gpurun_out
echo "=== kernels"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "exit $?"; tail -n 5 gpurun_out/pytest_kernels.log
for r in 8 16; do
  echo "=== WSI_STREAM_RING=$r"
  WSI_STREAM_RING=$r WSI_CONV_TRACE=1 timeout 300 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/conv_trace_ring$r.log 2>&1; echo "exit $?"
  grep -E "iter 2|BK16" gpurun_out/conv_trace_ring$r.log | cut -c1-110
done
