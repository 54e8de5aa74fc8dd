#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_v28.log
timeout 900 python bench.py --workload kitti --steps 2 --warmup 1 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_v28_kitti.json
timeout 900 python bench.py --workload 1080p --steps 2 --warmup 1 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_v28_1080p.json
