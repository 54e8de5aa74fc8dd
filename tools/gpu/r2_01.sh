#!/bin/bash
# Round 2, GPU call 1: parity of the new kernels (window-record arm sums, fused same-axis passes, chunked WTA), first numbers.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_01
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${O}_smi.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee ${O}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 | tee ${O}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > ${O}_bench_cone.json; cat ${O}_bench_cone.json | cut -c1-1500
timeout 300 python bench.py --workload kitti --steps 1 --warmup 1 --no-cpu 2>&1 | tail -1 > ${O}_bench_kitti.json; cut -c1-600 ${O}_bench_kitti.json
timeout 300 python bench.py --workload 1080p --steps 1 --warmup 1 --no-cpu 2>&1 | tail -1 > ${O}_bench_1080p.json; cut -c1-600 ${O}_bench_1080p.json
# unfused / budget A-B on cone (device-resident only)
for kb in 40 75 100; do ADC_AGG_SMEM_KB=$kb timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('smem_kb', $kb, d['value'], d['aggregation'])"; done 2>&1 | tee ${O}_ab.log
# launch list of one short bench run + full captures of the new kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file ${O}_launches.csv python tools/profile_one.py 32 2 > /dev/null 2>&1
python tools/summarize_launches.py ${O}_launches.csv > ${O}_launches_summary.txt 2>&1; head -40 ${O}_launches_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_arm_sum|k_wta|k_scanline|k_cost_volume' -s 20 -c 14 -o ${O}_full python tools/profile_one.py 32 2 > ${O}_ncu.log 2>&1
ls -la gpurun_out | tail -12
