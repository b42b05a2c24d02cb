#!/bin/bash
# round-2 call E: first run of the v3 ray-march kernel (compositing folded into the layer-2 operand, issuer warps)
mkdir -p gpurun_out
echo "== smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3 | cut -c1-400
echo "== bench_raymarch"; timeout 200 python scripts/bench_raymarch.py 2> gpurun_out/bench_raymarch.err | tee gpurun_out/bench_raymarch_r2e.jsonl | cut -c1-400; tail -3 gpurun_out/bench_raymarch.err
echo "== renderer tests"; timeout 300 python -m pytest tests/test_gpu_renderer.py -x -q -m gpu 2>&1 | tail -15 | cut -c1-600
echo "== full-size + generator + speedup tests"; timeout 420 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_generator.py tests/test_gpu_speedup.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -12 | cut -c1-600
echo "== bench"; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r2e.err | tee gpurun_out/bench_r2e.json | cut -c1-2500; tail -3 gpurun_out/bench_r2e.err
