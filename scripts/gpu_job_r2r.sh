#!/bin/bash
mkdir -p gpurun_out
echo "== memcheck round-2 kernels"; timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python scripts/sanitize_round2.py > gpurun_out/r2r_memcheck.txt 2>&1; tail -6 gpurun_out/r2r_memcheck.txt | cut -c1-300
echo "== smoke"; timeout 150 python __graft_entry__.py smoke 2>&1 | grep -i "smoke\|error" | cut -c1-400
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | cut -c1-400
