"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck are 10-100x slower than a plain run):
two small pairs through Match, the batched entry point, the voting fallbacks' sizes and the render calls."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import adcensus_b200 as A
import adc_testlib as T

cases = [(97, 61, 24, 2, {}), (130, 70, 37, 3, {}), (80, 60, 32, 10, {"do_discontinuity_adjustment": 1}),
         (80, 60, 32, 31, {"min_disparity": 2, "max_disparity": 34}),
         (70, 44, 64, 4, {}),        # D = 64: eight quads per CTA (compile-time strides), 8 lanes per scanline
         (150, 40, 130, 12, {}),     # 16 lanes per scanline, padded disparity stride
         (600, 16, 12, 18, {})]      # a row cut into segments by the fused horizontal double pass
for (w, h, D, seed, over) in cases:
    left, right = T.synthetic_pair(w, h, D, seed)
    kw = dict(max_disparity=D); kw.update(over)
    eng = A.Engine(w, h, A.ADCensusOption(**kw), wave_pairs=2, lanes=2)
    a = eng.match(left, right)
    b = eng.match_batch(np.stack([left] * 5), np.stack([right] * 5))
    assert (b.view(np.uint32) == a.view(np.uint32)[None]).all()
    want = T.Oracle(w, h, T.default_option(**kw)).match(left, right)
    assert a.tobytes() == want.tobytes(), (w, h, D)
    g, j, mm = eng.render_disparity(a)
    c = eng.disparity_cloud(left, a)
    eng.close()
    print("ok", w, h, D, mm, c.shape, flush=True)
print("all ok")
