#!/bin/bash
# round-2 call T: final validation of the committed state (what the driver runs at round end)
mkdir -p gpurun_out
echo "== smoke"; timeout 150 python __graft_entry__.py smoke 2>&1 | grep -i "smoke\|error" | cut -c1-400
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | cut -c1-400
echo "== bench"; timeout 600 python bench.py 2>gpurun_out/bench_r2t.err | tee gpurun_out/bench_r2t.json | cut -c1-2600; tail -2 gpurun_out/bench_r2t.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-500
