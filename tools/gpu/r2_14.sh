#!/bin/bash
# Round 2, GPU call 14: compile-time scanline sizes, 128-bit WTA left tile; validation + kernel timings + bench.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_14
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -6 | tee ${O}_pytest.log
K=scanline_x,scanline_y,wta
for wlk in cone kitti 1080p; do timeout 200 python tools/kernel_ab.py $wlk 2>&1 | tail -1 | tee -a ${O}_ab.log; done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['outputs_bit_identical'], d['aggregation'])"
