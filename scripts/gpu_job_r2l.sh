#!/bin/bash
# round-2 call L (8 GPUs): bench.py weak scaling with the shared-memory frame transport, config 3 (gen_videos grid 2x2 seeds 0-255) and config 4 (256^3 grid) at 8 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "== bench --gpus 8"; timeout 400 $TR --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 2>gpurun_out/bench_n8.err | grep '^{' | tee gpurun_out/bench_n8.json | cut -c1-1200; tail -3 gpurun_out/bench_n8.err | cut -c1-300
echo "== config 4 at 8 GPUs"; timeout 200 $TR --master-port 29512 scripts/bench_voxel_dist.py 2>gpurun_out/voxel_n8.err | grep '^{' | tee gpurun_out/voxel_n8.json | cut -c1-700; tail -2 gpurun_out/voxel_n8.err | cut -c1-300
echo "== config 3 at 8 GPUs"; timeout 500 $TR --master-port 29513 scripts/bench_video.py --seeds 256 --chunk 2048 2>gpurun_out/video_n8.err | grep '^{' | tee gpurun_out/video_n8.json | cut -c1-900; tail -3 gpurun_out/video_n8.err | cut -c1-300
