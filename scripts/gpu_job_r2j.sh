#!/bin/bash
# round-2 call J: whole GPU suite on the product build, bench line, launch list and ncu captures of the step's kernels, config 3 at 1 GPU
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | cut -c1-500
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_r2j.err | tee gpurun_out/bench_r2j.json | cut -c1-400; tail -2 gpurun_out/bench_r2j.err
echo "== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 700 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200; wc -l gpurun_out/r2j_launches.csv
echo "== ncu step kernels"; timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"modconv_epilogue|upfirdn2d_cl|style_|raymarch_tc3|bias_act" -c 40 -o gpurun_out/r2j_step_kernels -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log | cut -c1-200
echo "== ncu sigma_tc"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:sigma_tc_kernel -c 1 -o gpurun_out/r2j_sigma_tc -f python scripts/bench_voxel_dist.py > gpurun_out/ncu_vox.log 2>&1; tail -2 gpurun_out/ncu_vox.log | cut -c1-300
echo "== bench_video 1 GPU (64 seeds)"; timeout 400 python scripts/bench_video.py --seeds 64 2>gpurun_out/bench_video1.err | tee gpurun_out/bench_video_n1.json | cut -c1-700; tail -2 gpurun_out/bench_video1.err
ls -la gpurun_out/*.ncu-rep
