#!/bin/bash
# scan_resize branch: parity tests, then the whole GPU suite, then two probes (scan_resize 2 at the headline tile)
mkdir -p gpurun_out
echo "=== resize tests"; timeout 900 python -m pytest tests/test_gpu_resize.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/pytest_resize.log 2>&1; echo "exit $?"
grep -E "resize|passed|failed|Error|error" gpurun_out/pytest_resize.log | cut -c1-260 | tail -n 30
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_all.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/pytest_all.log | cut -c1-200
echo "=== probe: 512 tiles from 1024 windows, stride 256 (same 4x coverage in network pixels)"
timeout 300 python tools/perf_probe.py 20000 512 256 unet 0 2 2>&1 | grep -E "iter 2|gather|stitch|conv|stem" | tail -6
echo "=== probe: same slide, no resize, stride 128"
timeout 300 python tools/perf_probe.py 20000 512 128 unet 0 1 2>&1 | grep -E "iter 2|gather|stitch|conv|stem" | tail -6
