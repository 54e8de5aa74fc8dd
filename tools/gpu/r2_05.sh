#!/bin/bash
# Round 2, GPU call 5: TMA-staged fused aggregation kernel, warp-level bulk ring in the scanline kernel.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_05
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee ${O}_smoke.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "stage_parity or cone_all or golden_cases or real_pairs or baseline_configs or alternate or limits" 2>&1 | tail -15 | tee ${O}_pytest.log
K=arm_sum_h,arm_sum2_v,arm_sum2_h,arm_sum_h_div,scanline_x,scanline_y,wta
ab() { echo "== $*" | tee -a ${O}_ab.log; env "$@" timeout 120 python tools/kernel_ab.py cone $K 2>&1 | tail -1 | tee -a ${O}_ab.log; }
ab A=0
ab ADC_AGG2_TMA=0 ADC_SO_BULK=0
ab ADC_AGG2T_QC_V=4 ADC_AGG2T_QC_H=4
ab ADC_AGG2T_QC_V=8 ADC_AGG2T_QC_H=8 ADC_AGG2T_SMEM_KB=130
ab ADC_AGG2T_QC_V=4 ADC_AGG2T_QC_H=4 ADC_AGG2T_SMEM_KB=72
ab ADC_AGG2T_QC_V=4 ADC_AGG2T_QC_H=4 ADC_AGG2T_SMEM_KB=54
for wlk in kitti 1080p; do echo "== $wlk" | tee -a ${O}_ab.log; timeout 200 python tools/kernel_ab.py $wlk 2>&1 | tail -1 | tee -a ${O}_ab.log; ADC_AGG2T_QC_V=4 ADC_AGG2T_QC_H=4 ADC_AGG2T_SMEM_KB=72 timeout 200 python tools/kernel_ab.py $wlk arm_sum2_v,arm_sum2_h 2>&1 | tail -1 | tee -a ${O}_ab.log; done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > ${O}_bench_cone.json; cut -c1-330 ${O}_bench_cone.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_arm_sum2|k_scanline' -s 3 -c 4 -o ${O}_full python tools/profile_one.py 32 2 > ${O}_ncu.log 2>&1
tail -3 ${O}_ncu.log
