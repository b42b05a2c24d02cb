#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | grep -i "smoke\|error" | cut -c1-400
echo "== bench_raymarch"; timeout 150 python scripts/bench_raymarch.py "--only=v3" 2>&1 | tail -5 | cut -c1-300
echo "== renderer tests coop st3"; IDE3D_TC_STAGES=3 timeout 200 python -m pytest tests/test_gpu_renderer.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-300
echo "== renderer tests default"; timeout 200 python -m pytest tests/test_gpu_renderer.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-300
