#!/bin/bash
# round-2 call D: ncu of the round-2 ray-march kernel and of the new filtered_lrelu kernel; filtered_lrelu tests after the barrier-alignment fix
mkdir -p gpurun_out
echo "== filtered_lrelu tests"; timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "filtered_lrelu" 2>&1 | tail -25 | cut -c1-800
echo "== ncu raymarch"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:raymarch_tc_kernel -c 1 -o gpurun_out/r2d_raymarch -f python scripts/bench_raymarch.py "--only=teams=2 ray_major=1" > gpurun_out/ncu_rm.log 2>&1; tail -3 gpurun_out/ncu_rm.log | cut -c1-300
echo "== ncu filtered_lrelu"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:filtered_lrelu_fused2 -c 2 -o gpurun_out/r2d_flrelu -f python scripts/bench_ops.py --only filtered_lrelu --no-ref > gpurun_out/ncu_fl.log 2>&1; tail -3 gpurun_out/ncu_fl.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
