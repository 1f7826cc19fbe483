#!/bin/bash
# end-of-round evidence after the late kernel changes: smoke, default bench (driver settings), reference arm, ncu launch list of one batch
mkdir -p gpurun_out
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench (driver settings)"; timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cut -c1-300 gpurun_out/bench.json; tail -n 2 gpurun_out/bench.err
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; cut -c1-300 gpurun_out/bench_ref.json
echo "=== ncu batch"; bash tools/gpu_ncu_r2.sh
