#!/bin/bash
# round-2 evidence: one whole batch of the bench workload (every kernel of the step) with the metrics of
# profiles/r02_kernels.csv: duration, tensor-pipe activity, DRAM bytes.  (one ncu pass per gpurun call)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-library --no-kernel-table"
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed"
timeout 600 $CMD > gpurun_out/plain_r2.log 2>&1 && \
timeout 1500 ncu --metrics $M --clock-control none -s 2200 -c 130 --csv --log-file gpurun_out/r02_batch.csv $CMD > gpurun_out/ncu_r2.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_r2.log; wc -l gpurun_out/r02_batch.csv
