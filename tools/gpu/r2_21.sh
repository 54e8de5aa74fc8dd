#!/bin/bash
# Round 2, GPU call 21: one mbarrier per TMA box in the fused aggregation kernel.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_21
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "stage_parity or cone_all or real_pairs or baseline_configs or alternate or loaded or limits" 2>&1 | tail -6 | tee ${O}_pytest.log
K=arm_sum_h,arm_sum2_v,arm_sum2_h,arm_sum_h_div
for wlk in cone kitti 1080p; do timeout 200 python tools/kernel_ab.py $wlk $K 2>&1 | tail -1 | tee -a ${O}_ab.log; done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['outputs_bit_identical'], d['aggregation'])" | tee ${O}_bench.log
