#!/bin/bash
timeout 280 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "epilogue or upsample2d_add or scaled_bias_act or channels_last_patch" > gpurun_out/sanitizer_ops.txt 2>&1
grep -n "=========" gpurun_out/sanitizer_ops.txt | head -40 | cut -c1-220
