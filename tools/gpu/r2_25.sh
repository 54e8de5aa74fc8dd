#!/bin/bash
# Round 2, GPU call 25: k_vote_push with 512 threads (half an SM), WTA left tile swizzle, wave / lane sweep.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2_25
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x 2>&1 | tail -4 | tee ${O}_pytest.log
ADC_VP=512 ADC_VP_SLOTS=20480 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "stage_parity or cone_all or real_pairs or alternate or loaded" 2>&1 | tail -4 | tee -a ${O}_pytest.log
for wlk in cone kitti 1080p; do timeout 200 python tools/kernel_ab.py $wlk wta 2>&1 | tail -1 | tee -a ${O}_ab.log; done
B() { timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['outputs_bit_identical'], d['single_pair']['match_ms_median_of_20'], d['config']['wave_pairs'], d['config']['lanes'])"; }
echo "== default" | tee -a ${O}_bench.log; B | tee -a ${O}_bench.log
echo "== ADC_VP=512 ADC_VP_SLOTS=20480" | tee -a ${O}_bench.log; ADC_VP=512 ADC_VP_SLOTS=20480 B | tee -a ${O}_bench.log
echo "== ADC_VP=512 ADC_VP_SLOTS=20480 lanes 5" | tee -a ${O}_bench.log; ADC_VP=512 ADC_VP_SLOTS=20480 B --lanes 5 | tee -a ${O}_bench.log
echo "== lanes 5" | tee -a ${O}_bench.log; B --lanes 5 | tee -a ${O}_bench.log
echo "== lanes 3" | tee -a ${O}_bench.log; B --lanes 3 | tee -a ${O}_bench.log
echo "== wave 24 lanes 5" | tee -a ${O}_bench.log; B --wave-pairs 24 --lanes 5 | tee -a ${O}_bench.log
