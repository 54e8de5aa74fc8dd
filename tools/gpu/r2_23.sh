#!/bin/bash
# Round 2, GPU call 23: cost kernel with 4x4 register blocks; k_vote_push streams flat forward entries, 4-byte adjacency, 4-slot derive.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_23
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x 2>&1 | tail -6 | tee ${O}_pytest.log
for wlk in cone kitti 1080p; do timeout 200 python tools/kernel_ab.py $wlk cost_volume,arm_sum_h,wta 2>&1 | tail -1 | tee -a ${O}_ab.log; done
timeout 200 python - <<'PY' 2>&1 | tail -8 | tee ${O}_vote.log
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, adcensus_b200 as A, adc_testlib as T
left, right = T.load_cone()
h, w, _ = left.shape
eng = A.Engine(w, h, A.ADCensusOption(max_disparity=64), lanes=1)
for i in range(3): eng.match(left, right)
c = list(eng.counters())
print("cone counters:", c)
print(f"vote_push single pair: total {c[5]} us = lists {c[4]} + derive {c[6]} + push {c[7]} + collect {c[8]} + rest; rounds {c[2]} derives {c[3]} changes {c[12]} adj {c[14]} slots {c[10]}+{c[11]}")
eng.close()
PY
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['outputs_bit_identical'], d.get('kernels',{}).get('cost_volume'))" | tee ${O}_bench.log
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py 2>&1 | tail -12 | tee ${O}_memcheck.log
