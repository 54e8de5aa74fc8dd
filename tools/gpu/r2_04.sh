#!/bin/bash
# Round 2, GPU call 4: full parity run; bulk-async scanline ring A/B; arm sums with tabulated reciprocals + L1 prefetches.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_04
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee ${O}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -15 | tee ${O}_pytest.log
K=arm_sum_h,arm_sum2_v,arm_sum2_h,arm_sum_h_div,scanline_x,scanline_y,wta
ab() { echo "== $*" | tee -a ${O}_ab.log; env "$@" timeout 120 python tools/kernel_ab.py cone $K 2>&1 | tail -1 | tee -a ${O}_ab.log; }
ab A=0
ab ADC_SO_BULK=0
ab ADC_AGG_SMEM_KB=50
for wlk in kitti 1080p; do echo "== $wlk" | tee -a ${O}_ab.log; timeout 200 python tools/kernel_ab.py $wlk 2>&1 | tail -1 | tee -a ${O}_ab.log; ADC_SO_BULK=0 timeout 200 python tools/kernel_ab.py $wlk scanline_x,scanline_y 2>&1 | tail -1 | tee -a ${O}_ab.log; done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > ${O}_bench_cone.json; cut -c1-330 ${O}_bench_cone.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_arm_sum2|k_scanline' -s 3 -c 4 -o ${O}_full python tools/profile_one.py 32 2 > ${O}_ncu.log 2>&1
tail -3 ${O}_ncu.log
