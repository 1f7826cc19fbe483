#!/bin/bash
mkdir -p gpurun_out
echo "=== postproc + multiproc tests"; timeout 900 python -m pytest tests/test_gpu_postproc.py tests/test_gpu_host.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_pp.log 2>&1; echo "exit $?"; tail -n 30 gpurun_out/pytest_pp.log | cut -c1-300
