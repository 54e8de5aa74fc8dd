#!/bin/bash
# Round 2, GPU call 15: cost kernel with lane = disparity; validation, timings, bench of the three workloads.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_15
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -6 | tee ${O}_pytest.log
for wlk in cone kitti 1080p; do timeout 200 python tools/kernel_ab.py $wlk 2>&1 | tail -1 | tee -a ${O}_ab.log; done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['outputs_bit_identical'], d['aggregation'])"
timeout 300 python bench.py --workload kitti --steps 1 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['outputs_bit_identical'], d['pipeline_hbm'])"
timeout 300 python bench.py --workload 1080p --steps 1 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['outputs_bit_identical'], d['pipeline_hbm'])"
