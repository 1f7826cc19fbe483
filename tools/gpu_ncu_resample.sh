#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/perf_probe.py 6000 512 256 unet 0 2 > gpurun_out/probe_rs.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"resample|gather_kernel" -c 12 --csv --log-file gpurun_out/ncu_resample.csv python tools/perf_probe.py 6000 512 256 unet 0 2 > gpurun_out/ncu_rs.log 2>&1
echo "exit $?"; tail -3 gpurun_out/probe_rs.log
