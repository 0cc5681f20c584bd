#!/bin/bash
# 8-GPU scaling run of the sharded scoring metric, launched the way the driver launches it
mkdir -p gpurun_out
N=${1:-8}
( time timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 \
    bench.py --gpus $N --steps 50 --warmup 5 ) > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "rc=$?"
tail -c 1500 gpurun_out/bench_n$N.log; tail -5 gpurun_out/bench_n$N.err
