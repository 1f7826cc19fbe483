#!/bin/bash
# end-of-round evidence: full GPU suite, smoke, default bench (20 steps like the driver), configs[4] sweep
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-200
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench (driver settings)"; timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cut -c1-300 gpurun_out/bench.json; tail -n 2 gpurun_out/bench.err
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; cut -c1-300 gpurun_out/bench_ref.json
echo "=== c5"; timeout 1500 python bench.py --config c5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "exit $?"; cut -c1-200 gpurun_out/bench_c5.json
echo "=== c4"; timeout 900 python bench.py --config c4 --steps 3 --warmup 1 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "exit $?"; cat gpurun_out/bench_c4.json | cut -c1-1500
echo "=== ncu batch"; bash tools/gpu_ncu_r2.sh
