#!/usr/bin/env python
"""Generates tests/golden/golden_*.npz from the UNMODIFIED reference (oracle/_ref built from
/root/reference by oracle/Makefile).  Run in the build container; the fixtures are committed so the
GPU box (no /root/reference) can check against the real reference's outputs.

For every case: sha256 of every tap after every stage (bit-exact pin for all intermediates,
including the cost volumes) plus the full arrays of the small per-pixel maps and the final map.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import adc_testlib as T  # noqa: E402

CASES = {
    # name: (source, W, H, option overrides, seed or crop)
    "cone_full": ("cone", None, None, {}, None),
    "cone_crop": ("cone", 140, 100, {"max_disparity": 32}, (150, 120)),
    "synth_a": ("synth", 97, 61, {"max_disparity": 24}, 2),
    "synth_b": ("synth", 130, 70, {"max_disparity": 37}, 3),
    "synth_opts": ("synth", 120, 90, {"max_disparity": 48, "lambda_ad": 7, "lambda_census": 20, "so_p1": 0.7,
                                       "so_p2": 2.5, "irv_ts": 10, "irv_th": 0.3, "lrcheck_thres": 0.5,
                                       "cross_L1": 20, "cross_L2": 9}, 11),
    "synth_disc": ("synth", 80, 60, {"max_disparity": 32, "do_discontinuity_adjustment": 1}, 10),
}
FULL_TAPS = {"ARMS", "SUPCNT_H", "SUPCNT_V", "DISP_L", "DISP_R", "MISMATCHES", "OCCLUSIONS", "CENSUS_L"}


def case_inputs(name):
    src, w, h, over, extra = CASES[name]
    opt = T.default_option(**over)
    if src == "cone":
        left, right = T.load_cone()
        if w is not None:
            left, right = T.crop_pair(left, right, extra[0], extra[1], w, h)
    else:
        left, right = T.synthetic_pair(w, h, opt.max_disparity - opt.min_disparity, extra)
    return left, right, opt


def main():
    assert T.have_ref() or (T.build_oracle() or T.have_ref()), "oracle/_ref is required (build container only)"
    out_dir = T.GOLDEN_DIR
    for name in CASES:
        left, right, opt = case_inputs(name)
        h, w, _ = left.shape
        ref = T.Reference(w, h, opt)
        ref.begin(left, right)
        arrays, hashes = {}, {}
        for st in T.STAGES:
            ref.step()
            for tap in T.STAGE_TAPS[st]:
                a = ref.tap(tap)
                hashes[f"{st}/{tap}"] = T.sha(a)
                if tap in FULL_TAPS and (name != "cone_full" or (st in ("WTA", "MEDIAN") and tap in ("DISP_L", "DISP_R"))):
                    arrays[f"{st}__{tap}"] = a.copy()
        stock = ref.stock_match(left, right)
        assert T.sha(stock) == hashes["MEDIAN/DISP_L"], "staged runner and stock Match disagree"
        np.savez_compressed(out_dir / f"golden_{name}.npz", hashes=json.dumps(hashes), **arrays)
        print(name, w, h, "final sha", hashes["MEDIAN/DISP_L"][:16], "arrays", len(arrays))


if __name__ == "__main__":
    main()
