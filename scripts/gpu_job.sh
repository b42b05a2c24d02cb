#!/bin/bash
set -x
timeout 900 python bench.py 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_r01_n1.json; cut -c1-300 gpurun_out/bench_r01_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_r01_ref.json; cut -c1-400 gpurun_out/bench_r01_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file gpurun_out/launches_r1_v2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:raymarch_tc -s 1 -c 1 -o gpurun_out/raymarch_tc_r1_v2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rm.log 2>&1; tail -1 gpurun_out/ncu_rm.log | cut -c1-120
