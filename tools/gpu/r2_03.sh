#!/bin/bash
# Round 2, GPU call 3: leaner arm-sum kernels (no tail trip, 2-D thread blocks, line prefetch, immediate smem offsets).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_03
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "stage_parity or cone_all or golden_cases or real_pairs or baseline_configs or alternate" 2>&1 | tail -5 | tee ${O}_pytest.log
K=arm_sum_h,arm_sum2_v,arm_sum2_h,arm_sum_h_div,wta
ab() { echo "== $*" | tee -a ${O}_ab.log; env "$@" timeout 120 python tools/kernel_ab.py cone $K 2>&1 | tail -1 | tee -a ${O}_ab.log; }
ab A=0
ab ADC_ARM_PF=0
ab ADC_ARM_PF=296
ab ADC_ARM_PF=1184
ab ADC_AGG_SMEM_KB=50
for wlk in kitti 1080p; do echo "== $wlk" | tee -a ${O}_ab.log; timeout 200 python tools/kernel_ab.py $wlk 2>&1 | tail -1 | tee -a ${O}_ab.log; done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > ${O}_bench_cone.json; cut -c1-330 ${O}_bench_cone.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_arm_sum' -s 5 -c 5 -o ${O}_full python tools/profile_one.py 32 2 > ${O}_ncu.log 2>&1
tail -3 ${O}_ncu.log
