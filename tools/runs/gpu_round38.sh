#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_v38.log
python __graft_entry__.py smoke 2>&1 | tail -1
for cfg in "16 4" "24 3" "32 3"; do set -- $cfg; timeout 200 python bench.py --workload kitti --steps 1 --warmup 1 --no-cpu --wave-pairs $1 --lanes $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kitti', d['config']['wave_pairs'], d['config']['lanes'], d['value'])"; done | tee gpurun_out/kitti_sweep_v38.log
for cfg in "8 4" "16 2" "6 4"; do set -- $cfg; timeout 200 python bench.py --workload 1080p --steps 1 --warmup 1 --no-cpu --wave-pairs $1 --lanes $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1080p', d['config']['wave_pairs'], d['config']['lanes'], d['value'])"; done | tee gpurun_out/p1080_sweep_v38.log
