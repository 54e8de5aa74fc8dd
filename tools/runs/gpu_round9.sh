#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_region_voting_bytes|k_median_wavefront|k_wta_tile|k_cost_volume|k_interpolate|k_cross_arms" -c 7 -o gpurun_out/full_v9 -f python tools/profile_one.py 16 1 > gpurun_out/full_v9.log 2>&1
tail -3 gpurun_out/full_v9.log
ncu -i gpurun_out/full_v9.ncu-rep --page raw --csv > gpurun_out/full_v9_raw.csv 2>/dev/null
ncu -i gpurun_out/full_v9.ncu-rep --page source --csv -k regex:k_region_voting_bytes > gpurun_out/full_v9_vote_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
