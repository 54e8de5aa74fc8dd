#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_v24.log
timeout 300 python tools/quick_bench.py 256 16x4,16x6,32x3,32x4,24x4 2>&1 | tee gpurun_out/quick_v24.log
ADC_WTA_MODE=0 timeout 300 python tools/quick_bench.py 256 16x6 2>&1 | grep -E "maps/s|wta" | tee gpurun_out/quick_v24_wta0.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v24.csv python tools/profile_one.py 16 2 > gpurun_out/prof24.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_v24.csv | tee gpurun_out/launches_v24.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wta_walk|k_vote_push" -c 4 -o gpurun_out/full_v24 -f python tools/profile_one.py 16 1 > gpurun_out/full_v24.log 2>&1
ncu -i gpurun_out/full_v24.ncu-rep --page raw --csv > gpurun_out/full_v24_raw.csv 2>/dev/null
