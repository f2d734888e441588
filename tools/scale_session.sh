#!/bin/bash
# tools/scale_session.sh N : sharded Griffin-Lim bench (config 4) on N GPUs of one box -> gpurun_out/bench_nN.log
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "exit $?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-300; tail -2 gpurun_out/bench_n$N.err
