#!/bin/bash
mkdir -p gpurun_out
echo "=== ingest tests"; timeout 600 python -m pytest tests/test_gpu_ingest.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/pytest_ingest.log 2>&1; echo "exit $?"; grep -E "max abs|passed|failed|Error|error|assert" gpurun_out/pytest_ingest.log | cut -c1-300 | tail -30
echo "=== all gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --deselect tests/test_gpu_ingest.py > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -n 12 gpurun_out/pytest_gpu.log | cut -c1-300
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
