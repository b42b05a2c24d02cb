#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cl.json | cut -c1-250
timeout 300 python scripts/bench_ops.py 2>&1 | grep "float32" | grep -i "upfirdn" | grep channels_last | cut -c1-250
