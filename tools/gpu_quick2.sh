#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel + e2e tests"; timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py tests/test_gpu_round2.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_q.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/pytest_q.log | cut -c1-300
echo "=== trace"; WSI_CONV_TRACE=1 timeout 300 python tools/perf_probe.py 8192 512 128 unet > gpurun_out/conv_trace.log 2>&1; grep -E "iter 2|conv  |stem  " gpurun_out/conv_trace.log | tail -3; grep -E "ms/launch" gpurun_out/conv_trace.log | cut -c1-110
