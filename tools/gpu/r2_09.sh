#!/bin/bash
# bisect the batch-only mismatch (bench flag false since call 5)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_09
chk() { echo "== $*" | tee -a ${O}_bisect.log; env "$@" timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], 'flag', d['outputs_bit_identical'], 'pageable', d['e2e_pageable']['bit_identical'])" 2>&1 | tee -a ${O}_bisect.log; }
chk A=0
chk ADC_SO_DBG=1
chk ADC_SO_DBG=2
chk ADC_SO_DBG=4
chk ADC_SO_DBG=7
