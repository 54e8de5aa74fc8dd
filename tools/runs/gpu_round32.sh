#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_v32.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke_v32.log
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_v32.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v32.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu_v32.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_v32.csv | tee gpurun_out/launches_v32.txt | head -30
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_arm_sum|k_scanline|k_cost_volume|k_wta_tile|k_vote_push|k_vote_scan|k_median|k_interpolate_fast|k_cross_arms|k_gray_census|k_so_records" -c 24 -o gpurun_out/full_v32 -f python tools/profile_one.py 32 1 > gpurun_out/full_v32.log 2>&1
ncu -i gpurun_out/full_v32.ncu-rep --page raw --csv > gpurun_out/full_v32_raw.csv 2>/dev/null
