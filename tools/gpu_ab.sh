#!/bin/bash
# A/B/C on ONE box, alternating, with the SM clock sampled during each run (boxes differ in how hard the power cap bites).
# Variants are prebuilt libraries ab/lib<NAME>.so (git-ignored, travel with the gpurun snapshot):
#   <edit or check out the sources of the variant>; python -m wsi_segmentation_pipeline_b200.build
#   (WSI_EXTRA_NVCC_FLAGS="-D..." for compile-time switches); cp wsi_segmentation_pipeline_b200/libwsi_b200.so ab/libA.so
# usage: gpu_ab.sh A B C   — capi loads the library named by WSI_B200_LIB
mkdir -p gpurun_out
for round in 1 2; do
for v in "$@"; do
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits -lms 250 > gpurun_out/clk_$v.txt 2>/dev/null &
  SMI=$!
  WSI_B200_LIB=$PWD/ab/lib$v.so timeout 300 python tools/perf_probe.py 20000 512 128 unet 2>&1 | grep -E "iter 2|conv " | tail -2 | tr '\n' ' '
  kill $SMI 2>/dev/null; wait $SMI 2>/dev/null
  python - <<PY
import statistics
rows=[l.split(',') for l in open('gpurun_out/clk_$v.txt') if ',' in l]
clk=[float(r[0]) for r in rows]; pw=[float(r[1]) for r in rows]
hot=[c for c,p in zip(clk,pw) if p>500]
print(" | variant $v round $round: sm clock median under load %s MHz (%d samples), power max %.0f W" % (statistics.median(hot) if hot else None, len(hot), max(pw) if pw else 0))
PY
done
done
