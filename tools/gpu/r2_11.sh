#!/bin/bash
# Round 2, GPU call 11: fused kernels with records / divisors staged in shared memory; TMA per axis; fused vs single passes.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_11
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "stage_parity or cone_all or real_pairs or baseline_configs or alternate or loaded" 2>&1 | tail -6 | tee ${O}_pytest.log
K=arm_sum_h,arm_sum_v,arm_sum_h_div,arm_sum_v_div,arm_sum2_v,arm_sum2_h
for wlk in cone kitti 1080p; do
  for t in 0 1 2 3; do echo "== $wlk TMA=$t" | tee -a ${O}_ab.log; ADC_AGG2_TMA=$t timeout 200 python tools/kernel_ab.py $wlk $K 2>&1 | tail -1 | tee -a ${O}_ab.log; done
done
echo "== cone TMA=3 QC4" | tee -a ${O}_ab.log; ADC_AGG2_TMA=3 ADC_AGG2T_QC_V=4 ADC_AGG2T_QC_H=4 timeout 200 python tools/kernel_ab.py cone $K 2>&1 | tail -1 | tee -a ${O}_ab.log
echo "== cone LDG QC4" | tee -a ${O}_ab.log; ADC_AGG2_TMA=0 ADC_AGG_QC_V=2 ADC_AGG_QC_H=2 timeout 200 python tools/kernel_ab.py cone $K 2>&1 | tail -1 | tee -a ${O}_ab.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['outputs_bit_identical'], d['aggregation'])"
