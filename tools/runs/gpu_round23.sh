#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_v23.log
timeout 300 python tools/quick_bench.py 256 16x4,16x6 2>&1 | tee gpurun_out/quick_v23.log
ADC_WTA_MODE=0 timeout 300 python tools/quick_bench.py 256 16x4 2>&1 | grep -E "maps/s|wta" | tee gpurun_out/quick_v23_wta0.log
ADC_ARM_3P=1 timeout 300 python tools/quick_bench.py 256 16x4 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_v23_3p.log
ADC_ARM_APV=6 timeout 300 python tools/quick_bench.py 256 16x4 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_v23_apv6.log
ADC_ARM_3P=1 ADC_ARM_APV=6 timeout 300 python tools/quick_bench.py 256 16x4 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_v23_3p_apv6.log
ADC_ARM_3P=1 timeout 600 python -m pytest tests -m gpu -x -q -k "stage_parity or cone_all" 2>&1 | tail -3 | tee gpurun_out/pytest_v23_3p.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v23.csv python tools/profile_one.py 16 2 > gpurun_out/prof23.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_v23.csv | tee gpurun_out/launches_v23.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wta_walk|k_vote_push|k_scanline|k_cost_volume|k_so_records|k_gray" -c 10 -o gpurun_out/full_v23 -f python tools/profile_one.py 16 1 > gpurun_out/full_v23.log 2>&1
ncu -i gpurun_out/full_v23.ncu-rep --page raw --csv > gpurun_out/full_v23_raw.csv 2>/dev/null
