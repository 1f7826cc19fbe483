#!/bin/bash
mkdir -p gpurun_out
echo "=== resize tests"; timeout 900 python -m pytest tests/test_gpu_resize.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/pytest_resize.log 2>&1; echo "exit $?"
grep -E "resize|passed|failed|Error|error" gpurun_out/pytest_resize.log | cut -c1-260 | tail -n 30
for r in 2 3; do
echo "=== probe: 512 tiles from $((512*r)) windows, stride $((128*r))"
timeout 300 python tools/perf_probe.py 20000 512 $((128*r)) unet 0 $r 2>&1 | grep -E "iter 2|gather|stitch|conv|stem" | tail -5
done
