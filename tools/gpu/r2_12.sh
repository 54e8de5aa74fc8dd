#!/bin/bash
# Round 2, GPU call 12 (8 GPUs): BASELINE.json configs[4] -- 4096 pairs of 1242x375x128 owned by rank 0, NCCL scatter ->
# Match on 8 B200 -> NCCL gather; the same for the Cone batch.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_12
nvidia-smi -L | tee ${O}_smi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload kitti --sharded --steps 2 --warmup 1 2>&1 | tail -1 > ${O}_sharded_kitti_8gpu.json; cut -c1-2600 ${O}_sharded_kitti_8gpu.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --sharded --steps 3 --warmup 2 2>&1 | tail -1 > ${O}_sharded_cone_8gpu.json; cut -c1-2000 ${O}_sharded_cone_8gpu.json
