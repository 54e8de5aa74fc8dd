#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for m in 1 0; do
  echo "=== ADC_VOTE_MODE=$m"
  ADC_VOTE_MODE=$m timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "Error|passed|failed|FAILED" | head -30 | tee gpurun_out/pytest_gpu_mode$m.log
done
ADC_VOTE_MODE=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v6.csv python tools/profile_one.py 16 2 > gpurun_out/prof6.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_v6.csv | head -24 | tee gpurun_out/launches_v6.txt
ADC_SWEEP_S=1 ADC_VOTE_MODE=1 timeout 600 python tools/quick_bench.py 256 16x3,32x2,16x4 2>&1 | tail -20 | tee gpurun_out/quick_bench_v6.log
