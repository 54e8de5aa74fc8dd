#!/bin/bash
# Round 2, GPU call 18: region voting with the batch-wide cell index instead of per-pair CSR adjacency lists.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_18
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -8 | tee ${O}_pytest.log
timeout 300 python tools/quick_bench.py 64 32x2 2>&1 | tail -4 | tee ${O}_quick.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['outputs_bit_identical'], d['single_pair'])" | tee ${O}_bench.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file /tmp/l.csv python tools/profile_one.py 32 2 > /dev/null 2>&1; python tools/summarize_launches.py /tmp/l.csv | head -14 | tee ${O}_launches.txt
