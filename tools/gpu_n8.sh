#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N"
timeout 900 $RUN --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit $?"; cut -c1-1800 gpurun_out/bench_n$N.json; tail -n 4 gpurun_out/bench_n$N.err | cut -c1-200
timeout 300 $RUN --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref exit $?"; cut -c1-200 gpurun_out/bench_ref_n$N.json
