#!/bin/bash
# push-based voting + line-walking arm sums: parity, then A/B timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_v19.log
timeout 300 python tools/quick_bench.py 256 16x4 2>&1 | tee gpurun_out/quick_v19.log
ADC_ARM_MODE=0 timeout 300 python tools/quick_bench.py 256 16x4 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_v19_arm0.log
ADC_VOTE_MODE=1 timeout 300 python tools/quick_bench.py 256 16x4 2>&1 | grep -E "maps/s|voting" | tee gpurun_out/quick_v19_vote1.log
ADC_ARM_PFD=0 timeout 300 python tools/quick_bench.py 256 16x4 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_v19_pf0.log
ADC_ARM_PFD=48 timeout 300 python tools/quick_bench.py 256 16x4 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_v19_pf48.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v19.csv python tools/profile_one.py 16 2 > gpurun_out/prof19.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_v19.csv | tee gpurun_out/launches_v19.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_arm_sum_line|k_vote_push|k_vote_init|k_median|k_interpolate|k_cross_arms" -c 12 -o gpurun_out/full_v19 -f python tools/profile_one.py 16 1 > gpurun_out/full_v19.log 2>&1
ncu -i gpurun_out/full_v19.ncu-rep --page raw --csv > gpurun_out/full_v19_raw.csv 2>/dev/null
