#!/bin/bash
mkdir -p gpurun_out
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_all.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/pytest_all.log | cut -c1-200
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench (driver settings)"; timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cut -c1-200 gpurun_out/bench.json; tail -n 2 gpurun_out/bench.err
echo "=== ncu batch"; bash tools/gpu_ncu_r2.sh
