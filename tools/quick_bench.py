"""Rough device-resident throughput of the Cone batch (development aid, not the contract bench)."""
import os, sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import adcensus_b200 as A
import adc_testlib as T

left, right = T.load_cone()
h, w, _ = left.shape
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfgs = [tuple(int(v) for v in c.split("x")) for c in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["16x3", "8x3", "32x2"])]
dl = torch.from_numpy(np.repeat(left[None], n, 0)).cuda()
dr = torch.from_numpy(np.repeat(right[None], n, 0)).cuda()
dd = torch.empty((n, h, w), dtype=torch.float32, device="cuda")
for S, lanes in cfgs:
    eng = A.Engine(w, h, A.ADCensusOption(), wave_pairs=S, lanes=lanes)
    st = torch.cuda.current_stream()
    for _ in range(2):
        eng.match_batch_device(n, dl.data_ptr(), dr.data_ptr(), dd.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    reps = 3
    for _ in range(reps):
        eng.match_batch_device(n, dl.data_ptr(), dr.data_ptr(), dd.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"S={S} lanes={lanes}: {n} pairs in {ms:.2f} ms -> {n / ms * 1000:.1f} maps/s", flush=True)
    if (S, lanes) == cfgs[0]:
        for name in ("cost_volume", "arm_sum_h", "arm_sum2_v", "arm_sum2_h", "arm_sum_h_div", "scanline_x", "scanline_y", "wta"):
            kms, kb = eng.profile_kernel(name, 5)
            print(f"   kernel {name:14s} {kms*1000:8.1f} us per wave of {S}  -> {kb/kms/1e6:8.1f} GB/s algorithmic")
    eng.close()
if os.environ.get("ADC_SWEEP_S"):
    for S in (1, 2, 4, 8):
        eng = A.Engine(w, h, A.ADCensusOption(), wave_pairs=S, lanes=1)
        eng.match_batch_device(S, dl.data_ptr(), dr.data_ptr(), dd.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        line = f"S={S}: per-pair us:"
        for name in ("cost_volume", "arm_sum_h", "arm_sum2_v", "arm_sum2_h", "arm_sum_h_div", "scanline_x", "scanline_y", "wta"):
            kms, kb = eng.profile_kernel(name, 10)
            line += f" {name}={kms*1000/S:.1f}"
        print(line, flush=True)
        eng.close()
eng = A.Engine(w, h, A.ADCensusOption())
for _ in range(3):
    d = eng.match(left, right)
c = eng.counters()
print("voting counters [mism, occl, rounds, derives]:", c[:4], "list-build us:", c[4], "[changes, adjacency used, adjacency entries, forward-list room]:", c[12:16])
print("single-pair stage ms (cost, aggr, so, wta, refine, out):", [round(x, 3) for x in eng.last_stage_ms()])
