#!/bin/bash
# final regression on a fresh box: smoke(), full pytest -m gpu, default bench, reference arm (bounded)
mkdir -p gpurun_out
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -n 6 gpurun_out/smoke.log | cut -c1-200
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/pytest_gpu.log
echo "=== bench (defaults)"; timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "exit $?"; wc -l gpurun_out/bench_default.json; cut -c1-260 gpurun_out/bench_default.json
echo "=== bench --impl reference"; timeout 1200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; wc -l gpurun_out/bench_ref.json; cut -c1-300 gpurun_out/bench_ref.json
