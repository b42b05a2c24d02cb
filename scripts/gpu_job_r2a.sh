#!/bin/bash
# round-2 call A: diagnostics on the round-1 kernels (parity at the timed configuration, determinism, reference plugins, ncu of the in-step kernels)
mkdir -p gpurun_out
echo "== full-size parity"; timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -s -m gpu 2>&1 | grep -v "^$" | tail -15 | cut -c1-600
echo "== determinism small"; timeout 300 python scripts/determinism.py --small > gpurun_out/determinism_small.json 2> gpurun_out/determinism_small.err; tail -40 gpurun_out/determinism_small.json
echo "== determinism full"; timeout 300 python scripts/determinism.py > gpurun_out/determinism_full.json 2> gpurun_out/determinism_full.err; tail -40 gpurun_out/determinism_full.json
echo "== racecheck"; timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/racecheck_smoke.txt 2>&1; tail -15 gpurun_out/racecheck_smoke.txt | cut -c1-300
echo "== synccheck"; timeout 300 compute-sanitizer --tool synccheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/synccheck_smoke.txt 2>&1; tail -8 gpurun_out/synccheck_smoke.txt | cut -c1-300
echo "== bench_ops with reference plugins"; timeout 600 python scripts/bench_ops.py > gpurun_out/bench_ops_r2a.txt 2> gpurun_out/bench_ops_r2a.err; cat gpurun_out/bench_ops_r2a.txt | cut -c1-700; tail -3 gpurun_out/bench_ops_r2a.err
echo "== ncu in-step kernels"; timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"modconv_epilogue|upfirdn2d_cl" -c 45 -o gpurun_out/r2a_step_kernels -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1; tail -3 gpurun_out/ncu_step.log | cut -c1-300
echo "== all gpu tests"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
