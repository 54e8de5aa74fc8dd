#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_v20.log
timeout 300 python tools/quick_bench.py 256 16x4 2>&1 | tee gpurun_out/quick_v20.log
ADC_ARM_AP=2 timeout 300 python tools/quick_bench.py 256 16x4 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_v20_ap2.log
ADC_ARM_AP=6 timeout 300 python tools/quick_bench.py 256 16x4 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_v20_ap6.log
timeout 300 python tools/quick_bench.py 256 16x6,16x8,8x8 2>&1 | grep -E "maps/s" | tee gpurun_out/quick_v20_lanes.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v20.csv python tools/profile_one.py 16 2 > gpurun_out/prof20.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_v20.csv | tee gpurun_out/launches_v20.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_arm_sum|k_vote_push|k_vote_init|k_interpolate|k_cross_arms" -c 12 -o gpurun_out/full_v20 -f python tools/profile_one.py 16 1 > gpurun_out/full_v20.log 2>&1
ncu -i gpurun_out/full_v20.ncu-rep --page raw --csv > gpurun_out/full_v20_raw.csv 2>/dev/null
