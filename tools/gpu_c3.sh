#!/bin/bash
# BASELINE configs[2]: 100k x 80k slide, strong scaling.  usage: gpu_c3.sh N
N=${1:-1}
mkdir -p gpurun_out
free -g | head -2; nproc
if [ "$N" == "1" ]; then RUN="python bench.py"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py"; fi
timeout 1500 $RUN --gpus $N --config c3 --steps 2 --warmup 1 --no-library --no-kernel-table > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err; echo "exit $?"
cut -c1-2500 gpurun_out/bench_c3_n$N.json; tail -n 5 gpurun_out/bench_c3_n$N.err
