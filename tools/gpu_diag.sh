#!/bin/bash
# First-contact GPU diagnostics: each group in its own process (a trapped kernel poisons the CUDA
# context of its process only).  Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout 600 python -m pytest "$@" -q --no-header -p no:cacheprovider > gpurun_out/diag_$name.log 2>&1; echo "exit $?"; tail -n 15 gpurun_out/diag_$name.log; }
run hbm tests/test_gpu_kernels.py -k "gather or synth or maxpool"
run conv0 "tests/test_gpu_kernels.py::test_conv_igemm[n2_16x16_c64to64_k3s1]"
run conv tests/test_gpu_kernels.py -k "test_conv_igemm"
run up tests/test_gpu_kernels.py -k "upsample"
run stem tests/test_gpu_kernels.py -k "stem"
run e2e tests/test_gpu_e2e.py
echo "=== perf"; timeout 600 python tools/perf_probe.py > gpurun_out/perf_probe.log 2>&1; echo "exit $?"; tail -n 30 gpurun_out/perf_probe.log
