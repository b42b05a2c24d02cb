#!/bin/bash
mkdir -p gpurun_out
echo "== generator + fullsize + host tests"; timeout 600 python -m pytest tests/test_gpu_generator.py tests/test_gpu_fullsize.py tests/test_render_grad.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-400
echo "== bench"; timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r2s.err | tee gpurun_out/bench_r2s.json | cut -c1-260; tail -2 gpurun_out/bench_r2s.err
echo "== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 700 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; wc -l gpurun_out/r2s_launches.csv
