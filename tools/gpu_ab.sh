#!/bin/bash
mkdir -p gpurun_out
for b in 16 32 48 64 96; do
echo "=== batch $b"; timeout 600 python tools/perf_probe.py 6144 512 128 unet $b > gpurun_out/probe_b$b.log 2>&1; echo "exit $?"; grep -E "iter 2" gpurun_out/probe_b$b.log
done
