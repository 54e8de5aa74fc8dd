#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --print-limit 30 python tools/sanitize_small.py > gpurun_out/memcheck_v37.log 2>&1; tail -15 gpurun_out/memcheck_v37.log
timeout 400 compute-sanitizer --tool racecheck --print-limit 30 python tools/sanitize_small.py > gpurun_out/racecheck_v37.log 2>&1; tail -15 gpurun_out/racecheck_v37.log
