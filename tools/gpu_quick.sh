#!/bin/bash
mkdir -p gpurun_out
echo "=== kernels"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "exit $?"; tail -n 12 gpurun_out/pytest_kernels.log
echo "=== e2e"; timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/pytest_e2e.log 2>&1; echo "exit $?"; grep -E "rel err|vs bf16|passed|failed|FAILED|Error" gpurun_out/pytest_e2e.log | cut -c1-300 | tail -30
echo "=== conv trace"; WSI_CONV_TRACE=1 timeout 600 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/conv_trace.log 2>&1; echo "exit $?"; tail -n 42 gpurun_out/conv_trace.log
