#!/bin/bash
# same library, two runtime settings, one box, alternating: usage gpu_ab_env.sh "ENV=1 ..." [pytest -k expression]
mkdir -p gpurun_out
SETTING="$1"
if [ -n "$2" ]; then
  echo "=== correctness under: $SETTING"
  env $SETTING timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x -k "$2" > gpurun_out/pytest_env.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/pytest_env.log | cut -c1-200
fi
for round in 1 2; do
for v in base exp; do
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits -lms 250 > gpurun_out/clk_$v.txt 2>/dev/null &
  SMI=$!
  if [ $v == exp ]; then E="$SETTING"; else E="WSI_NOP=1"; fi
  env $E timeout 300 python tools/perf_probe.py 20000 512 128 unet 2>&1 | grep -E "iter 2|conv " | tail -2 | tr '\n' ' '
  kill $SMI 2>/dev/null; wait $SMI 2>/dev/null
  python - <<PY
import statistics
rows=[l.split(',') for l in open('gpurun_out/clk_$v.txt') if ',' in l]
clk=[float(r[0]) for r in rows]; pw=[float(r[1]) for r in rows]
hot=[c for c,p in zip(clk,pw) if p>500]
print(" | $v round $round: sm clock median under load %s MHz" % (statistics.median(hot) if hot else None))
PY
done
done
