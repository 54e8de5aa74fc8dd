#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_v30.log
timeout 300 python tools/quick_bench.py 256 32x4 2>&1 | tee gpurun_out/quick_v30.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v30.csv python tools/profile_one.py 32 2 > gpurun_out/prof30.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_v30.csv | head -16 | tee gpurun_out/launches_v30.txt
