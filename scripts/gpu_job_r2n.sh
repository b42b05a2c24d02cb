#!/bin/bash
T=ide-3d_b200/lib/libide3d_b200_tuning.so
for i in 1 2; do timeout 60 python scripts/debug_smoke.py 0 3 2>&1 | tail -1 | cut -c1-400; done
timeout 60 python scripts/debug_smoke.py 1 3 2>&1 | tail -1 | cut -c1-400
IDE3D_B200_LIB=$T IDE3D_TC_V2=1 timeout 60 python scripts/debug_smoke.py 0 3 2>&1 | tail -1 | cut -c1-400
IDE3D_B200_LIB=$T IDE3D_TC_COOP=0 IDE3D_TC_STAGES=2 timeout 60 python scripts/debug_smoke.py 0 3 2>&1 | tail -1 | cut -c1-400
IDE3D_B200_LIB=$T IDE3D_TC_COOP=1 IDE3D_TC_STAGES=2 timeout 60 python scripts/debug_smoke.py 0 3 2>&1 | tail -1 | cut -c1-400
IDE3D_B200_LIB=$T IDE3D_TC_COOP=1 IDE3D_TC_STAGES=3 timeout 60 python scripts/debug_smoke.py 0 3 2>&1 | tail -1 | cut -c1-400
IDE3D_B200_LIB=$T IDE3D_TC_COOP=1 IDE3D_TC_STAGES=3 timeout 60 python scripts/debug_smoke.py 1 3 2>&1 | tail -1 | cut -c1-400
