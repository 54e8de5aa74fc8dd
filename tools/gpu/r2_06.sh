#!/bin/bash
# Round 2, GPU call 6 (2 GPUs): scanline producers with incremental pointers, TMA on the horizontal double pass only;
# NCCL scatter / gather with the real engine (world size 2), sharded bench line at 2 GPUs.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_06
nvidia-smi -L | tee ${O}_smi.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "sharded or stage_parity or cone_all or real_pairs or baseline_configs or alternate" 2>&1 | tail -8 | tee ${O}_pytest.log
K=arm_sum_h,arm_sum2_v,arm_sum2_h,arm_sum_h_div,scanline_x,scanline_y,wta
for wlk in cone kitti 1080p; do timeout 200 python tools/kernel_ab.py $wlk 2>&1 | tail -1 | tee -a ${O}_ab.log; done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > ${O}_bench_cone.json; cut -c1-330 ${O}_bench_cone.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload kitti --sharded --steps 2 --warmup 1 2>&1 | tail -1 > ${O}_sharded_kitti_2gpu.json; cut -c1-1800 ${O}_sharded_kitti_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --sharded --steps 3 --warmup 2 2>&1 | tail -1 > ${O}_sharded_cone_2gpu.json; cut -c1-1500 ${O}_sharded_cone_2gpu.json
