#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "Error|passed|failed|FAILED" | head -30 | tee gpurun_out/pytest_gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v8.csv python tools/profile_one.py 16 2 > gpurun_out/prof8.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_v8.csv | head -24 | tee gpurun_out/launches_v8.txt
timeout 600 python tools/quick_bench.py 256 16x3,16x4,8x6 2>&1 | tail -12 | tee gpurun_out/quick_bench_v8.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_v8.log
