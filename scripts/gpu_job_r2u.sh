#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench_n2.err | grep '^{' | tee gpurun_out/bench_n2.json | cut -c1-900; tail -3 gpurun_out/bench_n2.err | cut -c1-300; df -h /dev/shm | tail -1
