#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_region_voting_bytes" -c 1 -o gpurun_out/full_v13 -f python tools/profile_one.py 16 1 > gpurun_out/full_v13.log 2>&1
ncu -i gpurun_out/full_v13.ncu-rep --page raw --csv > gpurun_out/full_v13_raw.csv 2>/dev/null
ncu -i gpurun_out/full_v13.ncu-rep --page source --csv > gpurun_out/full_v13_source.csv 2>/dev/null
ls -la gpurun_out | tail -4
