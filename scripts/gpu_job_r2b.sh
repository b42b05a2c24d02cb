#!/bin/bash
# round-2 call B: first run of the round-2 ray-march kernel (parity tests for both producer-team variants, timing of all variants, sanitizer)
mkdir -p gpurun_out
echo "== smoke"; timeout 180 python __graft_entry__.py smoke 2>&1 | tail -5 | cut -c1-400
echo "== renderer tests teams=3"; IDE3D_TC_TEAMS=3 timeout 300 python -m pytest tests/test_gpu_renderer.py -x -q -m gpu 2>&1 | tail -15 | cut -c1-400
echo "== renderer tests teams=2"; IDE3D_TC_TEAMS=2 timeout 300 python -m pytest tests/test_gpu_renderer.py -x -q -m gpu 2>&1 | tail -15 | cut -c1-400
echo "== bench_raymarch"; timeout 300 python scripts/bench_raymarch.py 2> gpurun_out/bench_raymarch.err | tee gpurun_out/bench_raymarch_r2b.jsonl | cut -c1-400; tail -3 gpurun_out/bench_raymarch.err
echo "== full-size + generator tests"; timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_generator.py tests/test_gpu_speedup.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -12 | cut -c1-500
echo "== memcheck smoke"; timeout 300 compute-sanitizer --tool memcheck --print-limit 10 python __graft_entry__.py smoke > gpurun_out/memcheck_smoke.txt 2>&1; tail -6 gpurun_out/memcheck_smoke.txt | cut -c1-300
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r2b.err | tee gpurun_out/bench_r2b.json | cut -c1-1500
