#!/bin/bash
# scratch A/B: pair kernel timing experiments (WSI_IGEMM_DBG; results are garbage for dbg != 0), auto batch 74
mkdir -p gpurun_out
for d in 0 1 2 3 4; do
  echo "=== WSI_IGEMM_DBG=$d"
  WSI_IGEMM_DBG=$d WSI_CONV_TRACE=1 timeout 300 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/conv_trace_pdbg$d.log 2>&1; echo "exit $?"
  grep -E "iter 2|cap=|BK64" gpurun_out/conv_trace_pdbg$d.log | cut -c1-112
done
echo "=== batch 64"
WSI_CONV_TRACE=1 timeout 300 python tools/perf_probe.py 4096 512 128 unet 64 > gpurun_out/conv_trace_b64.log 2>&1; echo "exit $?"
grep -E "iter 2|cap=" gpurun_out/conv_trace_b64.log | cut -c1-112
