#!/bin/bash
# one GPU visit: TMA probes, full GPU test-suite, TMA flavour of upfirdn2d, op micro-bench, step bench in both layouts
set -x
bash scripts/run_tma_probe.sh > gpurun_out/tma_probe2.txt 2>&1; cat gpurun_out/tma_probe2.txt
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
IDE3D_TMA=1 timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "upfirdn2d or conv2d_resample" 2>&1 | tail -5
timeout 300 python scripts/bench_ops.py > gpurun_out/bench_ops.txt 2>&1; grep -i "upfirdn\|upsample\|filter2d\|downsample" gpurun_out/bench_ops.txt | cut -c1-250
IDE3D_TMA=1 timeout 300 python scripts/bench_ops.py > gpurun_out/bench_ops_tma.txt 2>&1; grep -i "upfirdn\|upsample\|filter2d\|downsample" gpurun_out/bench_ops_tma.txt | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cl.json | cut -c1-600
IDE3D_CHANNELS_LAST=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_nchw.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file gpurun_out/launches_cl.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
