#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v3.csv python tools/profile_one.py 16 2 > gpurun_out/prof3.log 2>&1
tail -2 gpurun_out/prof3.log
python tools/summarize_launches.py gpurun_out/launches_v3.csv | tee gpurun_out/launches_v3.txt
timeout 600 python tools/quick_bench.py 128 2>&1 | tail -12 | tee gpurun_out/quick_bench.log
