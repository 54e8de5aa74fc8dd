#!/bin/bash
# Round 2, GPU call 13: compute-sanitizer on the new kernels, wave / lane sweeps, full validation, bench lines, profiles.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_13
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py 2>&1 | tail -25 > ${O}_memcheck.txt; tail -6 ${O}_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|all ok|^ok" | sort | uniq -c | sort -rn | head -20 > ${O}_racecheck.txt; head -12 ${O}_racecheck.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 | tee ${O}_pytest.log
sw() { wl=$1; shift; for cfg in "$@"; do set -- $cfg; timeout 200 python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --wave-pairs $1 --lanes $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', d['config']['wave_pairs'], d['config']['lanes'], d['value'], d['e2e']['value'], d['outputs_bit_identical'])"; done; }
sw cone "32 4" "32 5" "48 3" "40 4" "24 5" "32 6" 2>&1 | tee ${O}_sweep.log
sw kitti "32 4" "16 4" "24 4" "16 6" 2>&1 | tee -a ${O}_sweep.log
sw 1080p "12 3" "8 4" "6 4" "16 2" 2>&1 | tee -a ${O}_sweep.log
