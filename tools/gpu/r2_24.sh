#!/bin/bash
# Round 2, GPU call 24: cost kernel four rows per CTA, swizzled scanline ring (K = 8), conflict-free WTA right tile.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_24
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x 2>&1 | tail -6 | tee ${O}_pytest.log
for wlk in cone kitti 1080p; do timeout 200 python tools/kernel_ab.py $wlk cost_volume,scanline_x,scanline_y,wta 2>&1 | tail -1 | tee -a ${O}_ab.log; done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['outputs_bit_identical'], d['single_pair']['match_ms_median_of_20'])" | tee ${O}_bench.log
