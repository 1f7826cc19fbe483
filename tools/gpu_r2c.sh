#!/bin/bash
mkdir -p gpurun_out
echo "=== stitch tests"; timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q --no-header -p no:cacheprovider -k "stitch or canvas or unaligned" > gpurun_out/pytest_r2s.log 2>&1; echo "exit $?"; tail -n 5 gpurun_out/pytest_r2s.log | cut -c1-300
echo "=== bench"; timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cat gpurun_out/bench.json; tail -n 8 gpurun_out/bench.err
