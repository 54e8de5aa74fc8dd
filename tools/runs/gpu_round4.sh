#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for m in 1 0 2; do
  echo "=== ADC_VOTE_MODE=$m"
  ADC_VOTE_MODE=$m timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_mode$m.log
done
for m in 1 0; do
  ADC_VOTE_MODE=$m ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v5_mode$m.csv python tools/profile_one.py 16 2 > gpurun_out/prof5.log 2>&1
  python tools/summarize_launches.py gpurun_out/launches_v5_mode$m.csv | head -16 | tee gpurun_out/launches_v5_mode$m.txt
done
ADC_VOTE_MODE=1 timeout 600 python tools/quick_bench.py 256 16x3,8x3,32x2,16x4 2>&1 | tail -16 | tee gpurun_out/quick_bench_mode1.log
ADC_VOTE_MODE=0 timeout 600 python tools/quick_bench.py 256 16x3 2>&1 | tail -5 | tee gpurun_out/quick_bench_mode0.log
