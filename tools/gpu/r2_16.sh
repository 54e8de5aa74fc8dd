#!/bin/bash
# Round 2, GPU call 16: the numbers and captures that go into profiles/ -- bench lines of the three workloads (Cone with the
# CPU baseline), launch list of a bench run, full ncu capture of every kernel kind of one wave.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_16
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > ${O}_clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > ${O}_bench_cone.json; python -c "import json; d=json.load(open('${O}_bench_cone.json')); print(d['value'], d['e2e']['value'], d['outputs_bit_identical'], d['aggregation'], d['single_pair'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > ${O}_bench_reference.json; cut -c1-300 ${O}_bench_reference.json
timeout 300 python bench.py --workload kitti --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 > ${O}_bench_kitti.json; python -c "import json; d=json.load(open('${O}_bench_kitti.json')); print(d['value'], d['e2e']['value'], d['outputs_bit_identical'], d['pipeline_hbm'], d['aggregation'])"
timeout 300 python bench.py --workload 1080p --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 > ${O}_bench_1080p.json; python -c "import json; d=json.load(open('${O}_bench_1080p.json')); print(d['value'], d['e2e']['value'], d['outputs_bit_identical'], d['pipeline_hbm'], d['aggregation'])"
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file ${O}_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
python tools/summarize_launches.py ${O}_launches_bench.csv > ${O}_launches_bench_summary.txt 2>&1; head -12 ${O}_launches_bench_summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'^k_|k_arm|k_scanline|k_median|k_region' -s 40 -c 39 -o ${O}_full python tools/profile_one.py 32 2 > ${O}_ncu.log 2>&1
tail -2 ${O}_ncu.log
for wlk in kitti 1080p; do timeout 300 ncu --set full --clock-control none -k regex:'k_arm_sum|k_scanline|k_wta|k_cost' -s 13 -c 13 -o ${O}_full_$wlk python tools/kernel_ab.py $wlk cost_volume > ${O}_ncu_$wlk.log 2>&1; done
ls -la gpurun_out | tail -5
