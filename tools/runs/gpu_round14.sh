#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ADC_ARM_MODE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "stage_parity or cone_all" 2>&1 | grep -E "Error|passed|failed|FAILED" | head -10
ADC_ARM_MODE=1 timeout 600 python tools/quick_bench.py 256 16x3,16x4 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_bench_staged.log
ADC_ARM_MODE=0 timeout 600 python tools/quick_bench.py 256 16x4 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_bench_direct.log
