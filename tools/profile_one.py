"""One warm-up wave + one measured wave of Cone pairs on a single lane (for ncu launch lists)."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import adcensus_b200 as A
import adc_testlib as T

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
waves = int(sys.argv[2]) if len(sys.argv) > 2 else 2
left, right = T.load_cone()
h, w, _ = left.shape
n = S * waves
eng = A.Engine(w, h, A.ADCensusOption(), wave_pairs=S, lanes=1)
dl = torch.from_numpy(np.repeat(left[None], n, 0)).cuda()
dr = torch.from_numpy(np.repeat(right[None], n, 0)).cuda()
dd = torch.empty((n, h, w), dtype=torch.float32, device="cuda")
eng.match_batch_device(n, dl.data_ptr(), dr.data_ptr(), dd.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("launches", eng.launch_count)
