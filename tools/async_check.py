"""Quick check of the experimental async_refine lanes: a batch through an async engine must equal the plain engine's maps."""
import sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import adcensus_b200 as A
import adc_testlib as T
left, right = T.load_cone()
h, w, _ = left.shape
n = 24
ref = A.Engine(w, h, A.ADCensusOption(), wave_pairs=4, lanes=2)
want = ref.match(left, right); want2 = ref.match(right, left)
ref.close()
eng = A.Engine(w, h, A.ADCensusOption(), wave_pairs=4, lanes=2, async_refine=True)
L = np.stack([left if i % 2 == 0 else right for i in range(n)]); R = np.stack([right if i % 2 == 0 else left for i in range(n)])
dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
dd = torch.zeros((n, h, w), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream()
for rep in range(2):
    eng.match_batch_device(n, dl.data_ptr(), dr.data_ptr(), dd.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
out = dd.cpu().numpy()
ok = all(out[i].tobytes() == (want if i % 2 == 0 else want2).tobytes() for i in range(n))
print("async_refine batch identical:", ok, "config", eng.wave_pairs, eng.lanes)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.close()
for mode in (False, True):
    e = A.Engine(w, h, A.ADCensusOption(), async_refine=mode)
    dl2 = torch.from_numpy(np.repeat(left[None], 256, 0)).cuda(); dr2 = torch.from_numpy(np.repeat(right[None], 256, 0)).cuda()
    dd2 = torch.zeros((256, h, w), dtype=torch.float32, device="cuda")
    for _ in range(2): e.match_batch_device(256, dl2.data_ptr(), dr2.data_ptr(), dd2.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(3): e.match_batch_device(256, dl2.data_ptr(), dr2.data_ptr(), dd2.data_ptr(), st.cuda_stream)
    e1.record(st); torch.cuda.synchronize()
    print("async" if mode else "plain", e.wave_pairs, e.lanes, round(3 * 256 / e0.elapsed_time(e1) * 1000, 1), "maps/s", flush=True)
    e.close()
