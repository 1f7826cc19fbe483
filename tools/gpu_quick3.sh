#!/bin/bash
mkdir -p gpurun_out
echo "=== gather tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider -k "gather or golden or tiles_agree or stem" > gpurun_out/pytest_g.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/pytest_g.log | cut -c1-200
timeout 300 python tools/perf_probe.py 12000 512 128 unet 2>&1 | grep -E "iter 2|gather|stitch|maxpool|stem" | tail -5
