#!/bin/bash
# round-2 call C: ray-march v2 (teams=2 default; teams=3 without the register split) timing + tests, hierarchical tests, new filtered_lrelu
mkdir -p gpurun_out
echo "== smoke"; timeout 150 python __graft_entry__.py smoke 2>&1 | tail -3 | cut -c1-400
echo "== bench_raymarch"; timeout 240 python scripts/bench_raymarch.py 2> gpurun_out/bench_raymarch.err | tee gpurun_out/bench_raymarch_r2c.jsonl | cut -c1-400; tail -3 gpurun_out/bench_raymarch.err
echo "== renderer tests"; timeout 300 python -m pytest tests/test_gpu_renderer.py -x -q -m gpu 2>&1 | tail -15 | cut -c1-600
echo "== filtered_lrelu tests"; timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "filtered_lrelu" 2>&1 | tail -25 | cut -c1-800
echo "== filtered_lrelu bench"; timeout 300 python scripts/bench_ops.py --only filtered_lrelu 2> gpurun_out/bench_ops_fl.err | tee gpurun_out/bench_ops_fl_r2c.jsonl | cut -c1-700; tail -3 gpurun_out/bench_ops_fl.err
echo "== full-size + generator + speedup tests"; timeout 420 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_generator.py tests/test_gpu_speedup.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -12 | cut -c1-600
echo "== bench"; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r2c.err | tee gpurun_out/bench_r2c.json | cut -c1-1800; tail -3 gpurun_out/bench_r2c.err
