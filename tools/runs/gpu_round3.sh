#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v4.csv python tools/profile_one.py 16 2 > gpurun_out/prof4.log 2>&1
tail -2 gpurun_out/prof4.log
python tools/summarize_launches.py gpurun_out/launches_v4.csv | tee gpurun_out/launches_v4.txt
timeout 600 python tools/quick_bench.py 128 2>&1 | tail -12 | tee gpurun_out/quick_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_arm_sum|k_scanline|k_cost_volume|k_wta_right" -c 14 -o gpurun_out/full_v4 -f python tools/profile_one.py 16 1 > gpurun_out/full_v4.log 2>&1
tail -3 gpurun_out/full_v4.log
ls -la gpurun_out/
