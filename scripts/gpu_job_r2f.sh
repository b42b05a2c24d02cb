#!/bin/bash
# round-2 call F: smoke numbers, ncu of the v3 ray-march kernel, style plan test, bench
mkdir -p gpurun_out
echo "== smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | grep -i "smoke\|error" | cut -c1-400
echo "== style plan + generator tests"; timeout 300 python -m pytest tests/test_gpu_generator.py -x -q -m gpu 2>&1 | tail -12 | cut -c1-600
echo "== ncu raymarch v3"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:raymarch_tc3_kernel -c 1 -o gpurun_out/r2f_raymarch_v3 -f python scripts/bench_raymarch.py "--only=v3 teams=2" > gpurun_out/ncu_rm3.log 2>&1; tail -3 gpurun_out/ncu_rm3.log | cut -c1-300
echo "== bench"; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r2f.err | tee gpurun_out/bench_r2f.json | cut -c1-600; tail -3 gpurun_out/bench_r2f.err
