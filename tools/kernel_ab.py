"""Times individual pipeline kernels (adc_profile_kernel: CUDA events on the engine's stream, one wave) for A/B runs behind the
development switches (environment variables are read once per process, so every variant is its own process).
usage: kernel_ab.py [cone|kitti|1080p] kernel[,kernel...]"""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import adcensus_b200 as A
import adc_testlib as T

wl = sys.argv[1] if len(sys.argv) > 1 else "cone"
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["cost_volume", "arm_sum_h", "arm_sum2_v", "arm_sum2_h", "arm_sum_h_div", "scanline_x", "scanline_y", "wta"]
if wl == "cone":
    left, right = T.load_cone(); D = 64
else:
    w0, h0, D = (1242, 375, 128) if wl == "kitti" else (1920, 1080, 192)
    left, right = T.synthetic_pair(w0, h0, D, 1)
h, w, _ = left.shape
eng = A.Engine(w, h, A.ADCensusOption(max_disparity=D), lanes=1)
n = eng.wave_pairs
dl = torch.from_numpy(np.repeat(left[None], n, 0)).cuda()
dr = torch.from_numpy(np.repeat(right[None], n, 0)).cuda()
dd = torch.empty((n, h, w), dtype=torch.float32, device="cuda")
eng.match_batch_device(n, dl.data_ptr(), dr.data_ptr(), dd.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
out = []
for name in names:
    ms, by = eng.profile_kernel(name, 10)
    out.append(f"{name}={ms * 1000:.0f}us({by / ms / 1e6:.0f}GB/s)")
print(f"{wl} S={n}: " + " ".join(out), flush=True)
eng.close()
