#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ADC_ARM_MODE=4 timeout 900 python -m pytest tests -m gpu -x -q -k "stage_parity or cone_all or golden or kitti" 2>&1 | tail -5 | tee gpurun_out/pytest_v27_mode4.log
ADC_ARM_MODE=4 timeout 300 python tools/quick_bench.py 256 32x4 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_v27_mode4.log
timeout 300 python tools/quick_bench.py 256 32x4,32x6,32x8 2>&1 | grep -E "maps/s|arm_sum" | tee gpurun_out/quick_v27.log
ADC_ARM_MODE=4 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v27.csv python tools/profile_one.py 32 2 > gpurun_out/prof27.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_v27.csv | head -12 | tee gpurun_out/launches_v27.txt
ADC_ARM_MODE=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_arm_sum_staged_line" -c 4 -o gpurun_out/full_v27 -f python tools/profile_one.py 32 1 > gpurun_out/full_v27.log 2>&1
ncu -i gpurun_out/full_v27.ncu-rep --page raw --csv > gpurun_out/full_v27_raw.csv 2>/dev/null
