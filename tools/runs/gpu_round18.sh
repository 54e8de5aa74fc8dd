#!/bin/bash
# state check after session restore: parity, launch list, full captures, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_v18.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v18.csv python tools/profile_one.py 16 2 > gpurun_out/prof18.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_v18.csv | tee gpurun_out/launches_v18.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_scanline|k_arm_sum|k_cost_volume|k_wta_tile|k_region_voting|k_median" -c 14 -o gpurun_out/full_v18 -f python tools/profile_one.py 16 1 > gpurun_out/full_v18.log 2>&1
ncu -i gpurun_out/full_v18.ncu-rep --page raw --csv > gpurun_out/full_v18_raw.csv 2>/dev/null
timeout 600 python tools/quick_bench.py 256 16x4 2>&1 | tee gpurun_out/quick_v18.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_v18.log
