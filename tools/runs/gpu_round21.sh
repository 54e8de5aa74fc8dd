#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_v21.log
timeout 300 python tools/quick_bench.py 256 16x4,16x6,16x8 2>&1 | tee gpurun_out/quick_v21.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v21.csv python tools/profile_one.py 16 2 > gpurun_out/prof21.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_v21.csv | tee gpurun_out/launches_v21.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_arm_sum|k_vote_push|k_vote_regions|k_interpolate_fast|k_median" -c 14 -o gpurun_out/full_v21 -f python tools/profile_one.py 16 1 > gpurun_out/full_v21.log 2>&1
ncu -i gpurun_out/full_v21.ncu-rep --page raw --csv > gpurun_out/full_v21_raw.csv 2>/dev/null
