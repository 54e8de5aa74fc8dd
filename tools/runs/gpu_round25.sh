#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_v25.log
ADC_WTA_MODE=0 timeout 300 python tools/quick_bench.py 256 32x4,16x6 2>&1 | tee gpurun_out/quick_v25.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v25.csv python tools/profile_one.py 16 2 > gpurun_out/prof25.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_v25.csv | tee gpurun_out/launches_v25.txt
