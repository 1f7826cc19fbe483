#!/bin/bash
# tests + per-layer trace + bench + ncu evidence, one gpurun call
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/pytest_gpu.log
echo "=== conv trace"; WSI_CONV_TRACE=1 timeout 600 python tools/perf_probe.py 4096 512 128 unet > gpurun_out/conv_trace.log 2>&1; echo "exit $?"; tail -n 45 gpurun_out/conv_trace.log
echo "=== bench"; timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
echo "=== ncu"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_halo|conv_igemm" -s 36 -c 10 -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -n 5 gpurun_out/ncu_full.log
fi
