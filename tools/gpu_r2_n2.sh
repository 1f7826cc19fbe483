#!/bin/bash
# 2-GPU check: peer-store result + shared host e2e, vs the NCCL exchange
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-library --no-kernel-table"
echo "=== peer"; timeout 900 $RUN > gpurun_out/bench_n2_peer.json 2> gpurun_out/bench_n2_peer.err; echo "exit $?"; cut -c1-1500 gpurun_out/bench_n2_peer.json; tail -n 6 gpurun_out/bench_n2_peer.err
echo "=== nccl"; WSI_BENCH_NO_PEER=1 timeout 900 $RUN --no-e2e > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; echo "exit $?"; cut -c1-300 gpurun_out/bench_n2_nccl.json; tail -n 3 gpurun_out/bench_n2_nccl.err
