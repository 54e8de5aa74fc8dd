#!/bin/bash
# Round 2, GPU call 7: full parity run of the current build, bench lines of the three workloads, sharded-arm diagnostics at
# world size 1, launch list and full ncu captures of every kernel kind of one wave (for profiles/).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_07
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee ${O}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -12 | tee ${O}_pytest.log
for wlk in cone kitti 1080p; do timeout 200 python tools/kernel_ab.py $wlk 2>&1 | tail -1 | tee -a ${O}_ab.log; done
timeout 300 python bench.py --sharded --pairs 64 --steps 2 --warmup 1 2>&1 | tail -1 > ${O}_sharded_w1.json; cut -c1-2000 ${O}_sharded_w1.json
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > ${O}_bench_cone.json; cut -c1-400 ${O}_bench_cone.json
timeout 300 python bench.py --workload kitti --steps 1 --warmup 1 --no-cpu 2>&1 | tail -1 > ${O}_bench_kitti.json; cut -c1-200 ${O}_bench_kitti.json
timeout 300 python bench.py --workload 1080p --steps 1 --warmup 1 --no-cpu 2>&1 | tail -1 > ${O}_bench_1080p.json; cut -c1-200 ${O}_bench_1080p.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file ${O}_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
python tools/summarize_launches.py ${O}_launches_bench.csv > ${O}_launches_bench_summary.txt 2>&1; head -42 ${O}_launches_bench_summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'^k_|k_arm|k_scanline|k_median|k_region' -s 40 -c 39 -o ${O}_full python tools/profile_one.py 32 2 > ${O}_ncu.log 2>&1
tail -3 ${O}_ncu.log
