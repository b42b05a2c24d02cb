#!/bin/bash
mkdir -p gpurun_out
echo "== mesh tests"; timeout 300 python -m pytest tests/test_mesh.py -x -q -m gpu 2>&1 | tail -6 | cut -c1-500
echo "== filtered_lrelu tests"; timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "filtered_lrelu" 2>&1 | tail -5 | cut -c1-500
echo "== filtered_lrelu bench"; timeout 300 python scripts/bench_ops.py --only filtered_lrelu 2> gpurun_out/bench_ops_fl.err | tee gpurun_out/bench_ops_fl_r2m.jsonl | cut -c1-600; tail -2 gpurun_out/bench_ops_fl.err
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | cut -c1-500
echo "== smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | grep -i "smoke\|error" | cut -c1-300
