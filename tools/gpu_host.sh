#!/bin/bash
mkdir -p gpurun_out
echo "=== e2e (new)"; timeout 1200 python -m pytest tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider -s -k "config1 or stride_sweep or full_size" > gpurun_out/pytest_new.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/pytest_new.log | cut -c1-260
