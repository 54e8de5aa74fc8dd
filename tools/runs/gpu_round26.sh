#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_v26.log
timeout 300 python tools/quick_bench.py 256 32x4 2>&1 | tee gpurun_out/quick_v26.log
ADC_ARM_NV=2 timeout 300 python tools/quick_bench.py 256 32x4 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_v26_nv2.log
ADC_ARM_NV=2 timeout 600 python -m pytest tests -m gpu -x -q -k "stage_parity or cone_all" 2>&1 | tail -3 | tee gpurun_out/pytest_v26_nv2.log
ADC_ARM_AP=3 timeout 300 python tools/quick_bench.py 256 32x4 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_v26_ap3.log
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_v26.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v26.csv python tools/profile_one.py 32 2 > gpurun_out/prof26.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_v26.csv | tee gpurun_out/launches_v26.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_arm_sum|k_scanline|k_cost_volume|k_wta_tile|k_vote_push|k_vote_scan|k_median|k_interpolate_fast|k_cross_arms" -c 22 -o gpurun_out/full_v26 -f python tools/profile_one.py 32 1 > gpurun_out/full_v26.log 2>&1
ncu -i gpurun_out/full_v26.ncu-rep --page raw --csv > gpurun_out/full_v26_raw.csv 2>/dev/null
