#!/bin/bash
# scratch A/B on one box: programmatic dependent launch on / off (whole-iteration throughput)
mkdir -p gpurun_out
for v in "WSI_NONE=1" "WSI_NO_PDL=1" "WSI_NONE=1" "WSI_NO_PDL=1"; do
  echo "=== [$v]"
  env $v timeout 300 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/probe_ab.log 2>&1; echo "exit $?"
  grep -E "iter [12]|conv  |classes_hist" gpurun_out/probe_ab.log | cut -c1-100
done
