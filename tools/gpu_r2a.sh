#!/bin/bash
# round 2, first contact: new tests first (fast feedback), then the whole GPU suite, then a short bench
mkdir -p gpurun_out
echo "=== round-2 tests"; timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/pytest_r2.log 2>&1; echo "exit $?"
grep -E "rel err|prob max|74x512|passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_r2.log | cut -c1-400 | tail -60
echo "=== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --deselect tests/test_gpu_round2.py > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -n 15 gpurun_out/pytest_gpu.log | cut -c1-300
echo "=== bench"; timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
