#!/bin/bash
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
IDE3D_TMA=0 timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "upfirdn2d or conv2d_resample or filtered" 2>&1 | tail -3
timeout 300 python scripts/bench_ops.py > gpurun_out/bench_ops.txt 2>&1; grep -i "upfirdn\|upsample\|filter2d\|downsample\|lrelu" gpurun_out/bench_ops.txt | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cl.json | cut -c1-400
IDE3D_CHANNELS_LAST=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_nchw.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file gpurun_out/launches_cl.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
