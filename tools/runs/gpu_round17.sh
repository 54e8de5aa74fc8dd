#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v17.csv python tools/profile_one.py 16 2 > gpurun_out/prof17.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_v17.csv | tee gpurun_out/launches_v17.txt
timeout 900 ncu --set full --clock-control none -k regex:"k_scanline|k_arm_sum|k_cost_volume|k_wta_tile|k_cross_arms|k_interpolate|k_region_voting_bytes" -c 12 -o gpurun_out/full_v17 -f python tools/profile_one.py 16 1 > gpurun_out/full_v17.log 2>&1
ncu -i gpurun_out/full_v17.ncu-rep --page raw --csv > gpurun_out/full_v17_raw.csv 2>/dev/null
timeout 900 python bench.py --steps 3 --warmup 3 --lanes 4 2>&1 | tail -2 | tee gpurun_out/bench_v17.log
