#!/bin/bash
# Round 2, GPU call 26: final validation of the tree (all GPU tests, smoke) and the numbers / captures that go into profiles/
# (ncu reports stay on the box; only condensed CSVs travel back).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_26
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5 | tee ${O}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee ${O}_smoke.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 1000 > ${O}_clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > ${O}_bench_cone.json; python -c "import json; d=json.load(open('${O}_bench_cone.json')); print(d['value'], d['e2e']['value'], d['outputs_bit_identical'])"
timeout 300 python bench.py --workload kitti --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 > ${O}_bench_kitti.json
timeout 300 python bench.py --workload 1080p --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 > ${O}_bench_1080p.json
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file /tmp/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
python tools/summarize_launches.py /tmp/launches_bench.csv > ${O}_launches_bench_summary.txt 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'^k_|k_arm|k_scanline|k_median|k_region' -s 40 -c 39 -o /tmp/full_cone python tools/profile_one.py 32 2 > ${O}_ncu.log 2>&1
python tools/ncu_summary.py /tmp/full_cone.ncu-rep ${O}_ncu_full_cone_summary.csv
ncu -i /tmp/full_cone.ncu-rep --page raw --csv > ${O}_ncu_full_cone_raw.csv 2>/dev/null
ncu -i /tmp/full_cone.ncu-rep --page source --csv 2>/dev/null | gzip > ${O}_ncu_full_cone_source.csv.gz
for wlk in kitti 1080p; do timeout 400 ncu --set full --clock-control none -k regex:'k_arm_sum|k_scanline|k_wta|k_cost' -s 13 -c 13 -o /tmp/full_$wlk python tools/kernel_ab.py $wlk cost_volume > /dev/null 2>&1; python tools/ncu_summary.py /tmp/full_$wlk.ncu-rep ${O}_ncu_full_${wlk}_summary.csv; done
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py 2>&1 | tail -12 > ${O}_memcheck.txt; tail -3 ${O}_memcheck.txt
du -sh gpurun_out
