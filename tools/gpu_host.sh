#!/bin/bash
mkdir -p gpurun_out
echo "=== host"; timeout 900 python -m pytest tests/test_gpu_host.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/pytest_host.log 2>&1; echo "exit $?"; tail -n 15 gpurun_out/pytest_host.log | cut -c1-300
