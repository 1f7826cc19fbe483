#!/bin/bash
mkdir -p gpurun_out
echo "=== round-2 tests"; timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q --no-header -p no:cacheprovider -s "$@" > gpurun_out/pytest_r2.log 2>&1; echo "exit $?"
grep -E "rel err|prob max|74x512|passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_r2.log | cut -c1-400 | tail -70
