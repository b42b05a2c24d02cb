#!/bin/bash
mkdir -p gpurun_out
echo "== bench_raymarch"; timeout 200 python scripts/bench_raymarch.py "--only=v3" 2>&1 | tail -6 | cut -c1-300
echo "== renderer tests default"; timeout 200 python -m pytest tests/test_gpu_renderer.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-300
