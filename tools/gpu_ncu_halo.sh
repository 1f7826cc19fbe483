#!/bin/bash
# ncu --set full of the halo-resident pair kernels inside the bench workload
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 $CMD > gpurun_out/plain3.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_halo" -s 12 -c 12 -o gpurun_out/prof_halo $CMD > gpurun_out/ncu_halo.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_halo.log
