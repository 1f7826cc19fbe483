#!/bin/bash
mkdir -p gpurun_out
echo "=== ingest tests"; timeout 600 python -m pytest tests/test_gpu_ingest.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_ingest.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/pytest_ingest.log | cut -c1-200
echo "=== bench default"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cut -c1-900 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
echo "=== bench c1"; timeout 600 python bench.py --config c1 --steps 5 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "exit $?"; cut -c1-600 gpurun_out/bench_c1.json
echo "=== bench c4"; timeout 900 python bench.py --config c4 --steps 3 --warmup 1 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "exit $?"; cat gpurun_out/bench_c4.json; tail -n 3 gpurun_out/bench_c4.err
echo "=== bench c5"; timeout 1500 python bench.py --config c5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "exit $?"; cut -c1-600 gpurun_out/bench_c5.json; tail -n 3 gpurun_out/bench_c5.err
echo "=== ncu batch"; bash tools/gpu_ncu_r2.sh
