#!/bin/bash
T=ide-3d_b200/lib/libide3d_b200_tuning.so
IDE3D_DEBUG_DUMP=1 IDE3D_B200_LIB=$T IDE3D_TC_V2=1 timeout 90 python scripts/debug_smoke.py 0 1 2>&1 | tail -8 | cut -c1-900
IDE3D_DEBUG_DUMP=1 IDE3D_B200_LIB=$T IDE3D_TC_COOP=1 IDE3D_TC_STAGES=3 timeout 90 python scripts/debug_smoke.py 0 1 2>&1 | tail -8 | cut -c1-900
