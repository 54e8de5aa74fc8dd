#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_v29.log
timeout 300 python tools/quick_bench.py 256 32x4 2>&1 | tee gpurun_out/quick_v29.log
ADC_SO_LPS=16 timeout 300 python tools/quick_bench.py 256 32x4 2>&1 | grep -E "maps/s|scanline" | tee gpurun_out/quick_v29_lps16.log
ADC_SO_LPS=16 timeout 600 python -m pytest tests -m gpu -x -q -k "cone_all or golden" 2>&1 | tail -3 | tee gpurun_out/pytest_v29_lps16.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v29.csv python tools/profile_one.py 32 2 > gpurun_out/prof29.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_v29.csv | head -22 | tee gpurun_out/launches_v29.txt
