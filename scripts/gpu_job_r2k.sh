#!/bin/bash
# round-2 call K: backward kernel tests + timing; bench line; launch list; ncu captures exported to CSV on the box (gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== backward kernel tests"; timeout 400 python -m pytest tests/test_render_grad.py -x -q -m gpu 2>&1 | tail -12 | cut -c1-600
echo "== bench_backward"; timeout 300 python scripts/bench_backward.py 2>gpurun_out/bench_bwd.err | tee gpurun_out/bench_backward.json | cut -c1-700; tail -2 gpurun_out/bench_bwd.err
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_r2k.err | tee gpurun_out/bench_r2k.json | cut -c1-300; tail -2 gpurun_out/bench_r2k.err
echo "== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 700 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; wc -l gpurun_out/r2k_launches.csv
echo "== ncu raymarch_tc3 in the step"; timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:raymarch_tc3 -c 1 -o gpurun_out/r2k_raymarch_tc3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rm.log 2>&1; tail -1 gpurun_out/ncu_rm.log | cut -c1-200
echo "== ncu other step kernels"; timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"modconv_epilogue|upfirdn2d_cl|style_" -c 24 -o /tmp/r2k_step -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1; tail -1 gpurun_out/ncu_step.log | cut -c1-200
ncu -i /tmp/r2k_step.ncu-rep --page raw --csv > gpurun_out/r2k_step_kernels_raw.csv 2>/dev/null; ls -la gpurun_out/r2k_step_kernels_raw.csv
echo "== ncu sigma_tc"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:sigma_tc_kernel -c 1 -o gpurun_out/r2k_sigma_tc -f python scripts/bench_voxel_dist.py > gpurun_out/ncu_vox.log 2>&1; tail -1 gpurun_out/ncu_vox.log | cut -c1-300
echo "== ncu bwd"; timeout 300 ncu --set full --clock-control none -k regex:raymarch_bwd -c 1 -o /tmp/r2k_bwd -f python scripts/bench_backward.py 2 > gpurun_out/ncu_bwd.log 2>&1; ncu -i /tmp/r2k_bwd.ncu-rep --page raw --csv > gpurun_out/r2k_bwd_raw.csv 2>/dev/null; tail -1 gpurun_out/ncu_bwd.log | cut -c1-200
du -sh gpurun_out
