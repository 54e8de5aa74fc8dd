#!/bin/bash
# Round 2, GPU call 10: the scanline ring with its proxy fence -- full parity run, bench flags of the three workloads.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_10
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -12 | tee ${O}_pytest.log
for wlk in cone kitti 1080p; do timeout 200 python tools/kernel_ab.py $wlk scanline_x,scanline_y,wta 2>&1 | tail -1 | tee -a ${O}_ab.log; done
timeout 300 python bench.py --sharded --pairs 64 --steps 2 --warmup 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sharded w1', d['value'], d['outputs_bit_identical'], d['outputs_check'])"
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > ${O}_bench_cone.json; python -c "import json; d=json.load(open('${O}_bench_cone.json')); print(d['value'], d['outputs_bit_identical'], d['e2e']['value'], d['e2e_pageable'], d['single_pair'], d['aggregation'], d['roofline'], d.get('cpu_baseline'))"
timeout 300 python bench.py --workload kitti --steps 1 --warmup 1 --no-cpu 2>&1 | tail -1 > ${O}_bench_kitti.json; python -c "import json; d=json.load(open('${O}_bench_kitti.json')); print(d['value'], d['outputs_bit_identical'], d['pipeline_hbm'])"
timeout 300 python bench.py --workload 1080p --steps 1 --warmup 1 --no-cpu 2>&1 | tail -1 > ${O}_bench_1080p.json; python -c "import json; d=json.load(open('${O}_bench_1080p.json')); print(d['value'], d['outputs_bit_identical'], d['pipeline_hbm'])"
