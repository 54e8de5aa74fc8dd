#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_v35.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_v35_pipelined.json | cut -c1-400
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-pipeline 2>&1 | tail -1 | tee gpurun_out/bench_v35_nopipe.json | cut -c1-200
ADC_ARM_MINB=5 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_v35_minb5.json | cut -c1-200
