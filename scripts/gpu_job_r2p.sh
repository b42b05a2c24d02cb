#!/bin/bash
# round-2 call P: final validation of the committed state + ncu of ALL FIR / epilogue launches of one step (the b256 / b512 layers are the last ones)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== smoke"; timeout 150 python __graft_entry__.py smoke 2>&1 | grep -i "smoke\|error" | cut -c1-400
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | cut -c1-400
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_r2p.err | tee gpurun_out/bench_r2p.json | cut -c1-260; tail -2 gpurun_out/bench_r2p.err
echo "== ncu step FIR/epilogue kernels"; timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"modconv_epilogue_cl|upfirdn2d_cl" -c 70 -o /tmp/r2p_step -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1; tail -1 gpurun_out/ncu_step.log | cut -c1-200
ncu -i /tmp/r2p_step.ncu-rep --page raw --csv > gpurun_out/r2p_step_kernels_raw.csv 2>/dev/null; ls -la gpurun_out/r2p_step_kernels_raw.csv
