#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_generator.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -12 | cut -c1-400
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', d['value'], d['ms_per_step'], d['e2e']['value'])"
IDE3D_CONV1X1_MM=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('conv1x1 as matmul', d['value'], d['ms_per_step'], d['e2e']['value'])"
