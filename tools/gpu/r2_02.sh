#!/bin/bash
# Round 2, GPU call 2: what limits the fused aggregation kernel (ncu), A/B of its occupancy / prefetch / chunk switches, WTA v3.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_02
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "stage_parity or cone_all or golden_cases or real_pairs or baseline_configs" 2>&1 | tail -5 | tee ${O}_pytest.log
K=arm_sum_h,arm_sum2_v,arm_sum2_h,arm_sum_h_div,wta
ab() { echo "== $*" | tee -a ${O}_ab.log; env "$@" timeout 120 python tools/kernel_ab.py cone $K 2>&1 | tail -1 | tee -a ${O}_ab.log; }
ab A=0
ab ADC_AGG2_MINB=3
ab ADC_AGG2_PF=444
ab ADC_AGG2_PF=592
ab ADC_AGG2_PF=1184
ab ADC_AGG_QC_H=2 ADC_AGG_QC_V=2
ab ADC_AGG_QC_H=2 ADC_AGG_QC_V=2 ADC_AGG2_PF=1184
ab ADC_AGG_QC_H=1 ADC_AGG_QC_V=1 ADC_AGG2_PF=2368
ab ADC_AGG2_THREADS=128
ab ADC_AGG2_THREADS=128 ADC_AGG2_PF=592
ab ADC_ARM_PF=0
for wlk in kitti 1080p; do echo "== $wlk" | tee -a ${O}_ab.log; timeout 200 python tools/kernel_ab.py $wlk 2>&1 | tail -1 | tee -a ${O}_ab.log; ADC_AGG2_PF=592 timeout 200 python tools/kernel_ab.py $wlk arm_sum2_v,arm_sum2_h 2>&1 | tail -1 | tee -a ${O}_ab.log; done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > ${O}_bench_cone.json; cut -c1-330 ${O}_bench_cone.json
# full captures: one launch of each kernel kind of the second wave (profile_one 32 2: 78 launches, second wave = 39 onwards)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_arm_sum|k_wta|k_cost_volume' -s 7 -c 7 -o ${O}_full python tools/profile_one.py 32 2 > ${O}_ncu.log 2>&1
tail -3 ${O}_ncu.log
