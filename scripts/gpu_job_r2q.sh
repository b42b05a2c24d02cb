#!/bin/bash
mkdir -p gpurun_out
echo "== ops + generator tests"; timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_generator.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -4 | cut -c1-400
echo "== bench_ops upsample2d_add"; timeout 200 python scripts/bench_ops.py --only upsample2d_add --no-ref 2>/dev/null | cut -c1-400
echo "== bench"; timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r2q.err | tee gpurun_out/bench_r2q.json | cut -c1-260; tail -2 gpurun_out/bench_r2q.err
