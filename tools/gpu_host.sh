#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_host.log 2>&1; echo "exit $?"; tail -n 30 gpurun_out/pytest_host.log | cut -c1-220
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 8 gpurun_out/smoke.log
