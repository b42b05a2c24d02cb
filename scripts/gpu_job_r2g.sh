#!/bin/bash
mkdir -p gpurun_out
echo "== debug v2/v3 t2"; timeout 120 python scripts/debug_v3.py v2 "v3 t2" fp32 2>&1 | tail -14 | cut -c1-300
echo "== debug v3 t3"; timeout 60 python scripts/debug_v3.py "v3 t3" 2>&1 | tail -6 | cut -c1-300
echo "== bench_raymarch v3 t3"; IDE3D_TC_TEAMS=3 timeout 90 python scripts/bench_raymarch.py "--only=v3" 2>&1 | tail -4 | cut -c1-300
echo "== voxel tests"; timeout 200 python -m pytest tests/test_gpu_renderer.py -x -q -m gpu -k "voxel or sigma_grid" 2>&1 | tail -8 | cut -c1-500
echo "== voxel bench"; timeout 120 python scripts/bench_voxel_dist.py 2>&1 | tail -2 | cut -c1-600
echo "== style plan test"; timeout 200 python -m pytest tests/test_gpu_generator.py -x -q -m gpu -k style_plan 2>&1 | tail -5 | cut -c1-500
