#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ADC_ARM_MINB=5 timeout 300 python tools/quick_bench.py 256 32x4 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_v34_minb5.log
timeout 300 python tools/quick_bench.py 256 32x4 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_v34.log
ADC_ARM_MINB=5 timeout 600 python -m pytest tests -m gpu -x -q -k "cone_all or golden" 2>&1 | tail -3 | tee gpurun_out/pytest_v34_minb5.log
