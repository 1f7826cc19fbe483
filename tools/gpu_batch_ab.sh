#!/bin/bash
# tiles per forward batch: same box, alternating
for round in 1 2; do
for b in ${BATCHES:-74 148 222}; do
  timeout 300 python tools/perf_probe.py 20000 512 128 unet $b 2>&1 | grep -E "iter 2|conv " | tail -2 | tr '\n' ' '; echo " | batch $b"
done
done
