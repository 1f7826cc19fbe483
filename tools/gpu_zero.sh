#!/bin/bash
mkdir -p gpurun_out
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_all.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/pytest_all.log | cut -c1-200
echo "=== probe 20000"
timeout 300 python tools/perf_probe.py 20000 512 128 unet 2>&1 | grep -E "iter 2|gather|stitch|conv|stem|maxpool" | tail -7
