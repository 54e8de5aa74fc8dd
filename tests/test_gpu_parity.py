"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, stage by stage.

Bar (BASELINE.json north_star): bit-exact census / arms / support counts / WTA integer indices;
<= 1e-4 on float costs; <= 0.01 px on sub-pixel disparity.  The kernels are built to be bit-exact
on the float stages too, and these tests assert exact equality there as well (a failure prints the
max abs difference so a tolerance-level deviation can be told from a real bug).
"""
import numpy as np
import pytest

import adc_testlib as T

pytestmark = pytest.mark.gpu


def _engine(w, h, opt, **kw):
    import adcensus_b200 as A
    o = A.ADCensusOption()
    for name, _ in T.Option._fields_:
        if not name.startswith("_"):
            setattr(o, name, getattr(opt, name))
    return A.Engine(w, h, o, **kw)


def _same(name, got, want):
    assert got.shape == want.shape, f"{name}: shape {got.shape} vs {want.shape}"
    if got.dtype.kind == "f":
        eq = (got.view(np.uint32) == want.view(np.uint32))
        if not eq.all():
            fin = np.isfinite(got) & np.isfinite(want)
            md = float(np.abs(got[fin].astype(np.float64) - want[fin]).max()) if fin.any() else 0.0
            inf_mismatch = int((np.isfinite(got) != np.isfinite(want)).sum())
            raise AssertionError(f"{name}: {int((~eq).sum())} of {eq.size} values differ, max abs diff {md:.3e}, "
                                 f"{inf_mismatch} finite/inf mismatches")
    else:
        assert np.array_equal(got, want), f"{name}: {int((got != want).sum())} of {got.size} values differ"


CASES = [
    # (W, H, D, option overrides, seed)
    (64, 48, 16, {}, 1),
    (97, 61, 24, {}, 2),            # odd sizes
    (130, 70, 37, {}, 3),           # D not a multiple of 4 (padded stride)
    (50, 40, 64, {}, 4),            # D > W: out-of-image matches everywhere
    (9, 12, 8, {}, 5),              # W <= 9: census early return (adcensus_util.cpp:12)
    (40, 7, 8, {}, 6),              # H <= 7
    (80, 60, 32, {"cross_L1": 10, "cross_L2": 4, "cross_t1": 30, "cross_t2": 12, "so_tso": 25}, 7),
    (80, 60, 32, {"do_lr_check": 0}, 8),
    (80, 60, 32, {"do_filling": 0}, 9),
    (80, 60, 32, {"do_discontinuity_adjustment": 1}, 10),
    (120, 90, 48, {"lambda_ad": 7, "lambda_census": 20, "so_p1": 0.7, "so_p2": 2.5, "irv_ts": 10,
                   "irv_th": 0.3, "lrcheck_thres": 0.5}, 11),
    (150, 100, 130, {}, 12),        # 16 lanes per line, 9 values -> padded stride
    (80, 60, 32, {"min_disparity": 2, "max_disparity": 34}, 31),     # dmin > 0
    (80, 60, 32, {"min_disparity": -4, "max_disparity": 28}, 32),    # negative dmin
    (200, 40, 200, {}, 13),         # a whole warp per line
]


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}x{c[1]}x{c[2]}-{i}" for i, c in enumerate(CASES)])
def test_stage_parity_synthetic(case):
    w, h, D, over, seed = case
    opt = T.default_option(**{"max_disparity": D, **over})
    left, right = T.synthetic_pair(w, h, D, seed)
    orc = T.Oracle(w, h, opt)
    eng = _engine(w, h, opt)
    orc.begin(left, right)
    for st in T.STAGES:
        orc.step()
        eng.debug_run(left, right, st)
        for tap in T.STAGE_TAPS[st]:
            _same(f"{st}/{tap}", eng.tap(tap), orc.tap(tap))
    # the public entry point gives the same map as the staged run
    _same("match", eng.match(left, right), orc.tap("DISP_L"))
    eng.close()


def test_cone_all_stages(cone):
    left, right = cone
    h, w, _ = left.shape
    orc = T.Oracle(w, h)
    eng = _engine(w, h, T.default_option())
    orc.begin(left, right)
    for st in T.STAGES:
        orc.step()
        eng.debug_run(left, right, st)
        for tap in T.STAGE_TAPS[st]:
            _same(f"cone {st}/{tap}", eng.tap(tap), orc.tap(tap))
    final = eng.match(left, right)
    _same("cone match", final, orc.tap("DISP_L"))
    eng.close()


def test_batch_equals_single():
    w, h, D = 96, 64, 32
    opt = T.default_option(max_disparity=D)
    pairs = [T.synthetic_pair(w, h, D, 100 + i) for i in range(11)]
    lefts = np.stack([p[0] for p in pairs])
    rights = np.stack([p[1] for p in pairs])
    eng = _engine(w, h, opt, wave_pairs=4, lanes=2)
    singles = [eng.match(l, r) for l, r in pairs]
    batch = eng.match_batch(lefts, rights)
    for i in range(len(pairs)):
        _same(f"batch[{i}]", batch[i], singles[i])
    ptrs = eng.match_batch_ptrs([p[0] for p in pairs], [p[1] for p in pairs])
    for i in range(len(pairs)):
        _same(f"ptrs[{i}]", ptrs[i], singles[i])
    orc = T.Oracle(w, h, opt)
    _same("vs oracle", singles[3], orc.match(*pairs[3]))
    eng.close()


def test_error_truth_table():
    """Mirrors ADCensusStereo.cpp:31,38,71,74: bad sizes / empty range / null pointers -> false."""
    import adcensus_b200 as A
    s = A.ADCensusStereo()
    assert s.Match(np.zeros((4, 4, 3), np.uint8), np.zeros((4, 4, 3), np.uint8)) is False  # before Initialize
    assert s.Initialize(0, 10, A.ADCensusOption()) is False
    assert s.Initialize(10, -1, A.ADCensusOption()) is False
    assert s.Initialize(10, 10, A.ADCensusOption(min_disparity=5, max_disparity=5)) is False
    assert s.Initialize(32, 24, A.ADCensusOption(max_disparity=8)) is True
    assert s.Match(None, np.zeros((24, 32, 3), np.uint8)) is False
    out = s.Match(np.zeros((24, 32, 3), np.uint8), np.zeros((24, 32, 3), np.uint8))
    assert out.shape == (24, 32)
    assert s.Reset(40, 30, A.ADCensusOption(max_disparity=16)) is True
    assert s.Match(np.zeros((30, 40, 3), np.uint8), np.zeros((30, 40, 3), np.uint8)).shape == (30, 40)
    s.Release()


def test_golden_cases_on_gpu():
    """The committed golden vectors (produced by the unmodified reference, tools/make_golden.py):
    every tap after every stage must hash to the reference's sha256."""
    import json
    import sys
    sys.path.insert(0, str(T.REPO / "tools"))
    import make_golden as G
    for name in ("cone_crop", "synth_a", "synth_b", "synth_opts", "synth_disc"):
        left, right, opt = G.case_inputs(name)
        z = np.load(T.GOLDEN_DIR / f"golden_{name}.npz")
        hashes = json.loads(str(z["hashes"]))
        h, w, _ = left.shape
        eng = _engine(w, h, opt)
        for st in T.STAGES:
            eng.debug_run(left, right, st)
            for tap in T.STAGE_TAPS[st]:
                assert T.sha(eng.tap(tap)) == hashes[f"{st}/{tap}"], f"{name}: {st}/{tap}"
        eng.close()


def test_cone_final_equals_reference_golden(cone):
    import json
    left, right = cone
    z = np.load(T.GOLDEN_DIR / "golden_cone_full.npz")
    hashes = json.loads(str(z["hashes"]))
    h, w, _ = left.shape
    eng = _engine(w, h, T.default_option())
    out = eng.match(left, right)
    assert T.sha(out) == hashes["MEDIAN/DISP_L"] and hashes["MEDIAN/DISP_L"].startswith("77d70a58d1aa5c71")
    # batched + pinned-async entry points give the same bits
    import torch
    n = 20
    hl = torch.from_numpy(np.repeat(left[None], n, 0)).pin_memory()
    hr = torch.from_numpy(np.repeat(right[None], n, 0)).pin_memory()
    hd = torch.empty((n, h, w), dtype=torch.float32).pin_memory()
    eng.match_batch_pinned_async(n, hl.data_ptr(), hr.data_ptr(), hd.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (hd.numpy().view(np.uint32) == out.view(np.uint32)[None]).all()
    eng.close()


@pytest.mark.parametrize("shape", [(1242, 375, 128), (1920, 1080, 192)], ids=["kitti_shape", "1080p"])
def test_full_size_properties(shape):
    """BASELINE configs 3/4 shapes.  Size-independent properties (the oracle needs 16 s / 115 s per
    pair here): (1) the synthetic pair has a known band disparity -> the map must recover it away
    from band edges / borders; (2) batch == single, bit for bit; (3) determinism."""
    w, h, D = shape
    opt = T.default_option(max_disparity=D)
    left, right = T.synthetic_pair(w, h, D, 1)
    eng = _engine(w, h, opt)
    a = eng.match(left, right)
    b = eng.match(left, right)
    assert a.view(np.uint32).tobytes() == b.view(np.uint32).tobytes()
    batch = eng.match_batch(np.stack([left, left, left]), np.stack([right, right, right]))
    assert (batch.view(np.uint32) == a.view(np.uint32)[None]).all()
    # ground truth: right(x) = left(x + d_band)  ->  disparity d_band on rows of the band
    lo, span = D // 8, max(1, (3 * D) // 4 - D // 8)
    bands = T._splitmix64(np.uint64(1) * np.uint64(1000003) + (np.arange(h) // 25).astype(np.uint64))
    truth = (lo + (bands % np.uint64(span)).astype(np.int64)).astype(np.float32)[:, None]
    inner = np.zeros((h, w), bool)
    for y in range(h):
        if 6 <= y % 25 <= 18:
            inner[y, D:w - 8] = True
    err = np.abs(a - truth)
    good = (err[inner] <= 1.0).mean()
    assert good > 0.9, f"only {good:.3f} of interior pixels within 1 px of the synthetic ground truth"
    eng.close()


def test_kitti_shape_vs_oracle():
    """One full-size 1242x375x128 pair against the CPU oracle (about 16 s of CPU)."""
    w, h, D = 1242, 375, 128
    opt = T.default_option(max_disparity=D)
    left, right = T.synthetic_pair(w, h, D, 3)
    eng = _engine(w, h, opt)
    got = eng.match(left, right)
    want = T.Oracle(w, h, opt).match(left, right)
    _same("kitti-shape final map", got, want)
    eng.close()


def test_cpp_dropin_program_on_gpu(tmp_path):
    """The reference-style C++ caller (tests/cpp/dropin_main.cpp) really runs Match on the device."""
    import os
    import subprocess
    import adcensus_b200 as A
    exe = tmp_path / "dropin"
    r = subprocess.run(["g++", "-std=c++17", str(T.REPO / "tests" / "cpp" / "dropin_main.cpp"), f"-I{T.REPO / 'include'}",
                        f"-L{A.lib_path().parent}", "-ladcensus_b200", f"-Wl,-rpath,{A.lib_path().parent}", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0 and "DROPIN_OK" in run.stdout, (run.returncode, run.stdout[-500:], run.stderr[-500:])
    assert "cost aggregating! timing" in run.stdout     # the reference's six timing lines are kept


@pytest.mark.parametrize("env", [{"ADC_VOTE_ENUM": "1"}, {"ADC_VOTE_SLOTCAP": "256"}], ids=["enumeration", "global-state"])
def test_voting_fallback_paths(tmp_path, env):
    """(enumeration) The voting kernel has two ways to find the histograms a filled pixel belongs to: precomputed adjacency lists
    (default) and on-the-fly enumeration of the inverse cross region (when the lists would not fit).  The switch is
    read once per process, so the fallback runs in a child: Cone and two synthetic cases through the VOTE stage and
    the final map, against the oracle.
    (global-state) Per-slot state lives in shared memory when the lists fit, else in global memory; a tiny capacity forces
    the latter."""
    import os, subprocess, sys, textwrap
    script = tmp_path / "enum_case.py"
    script.write_text(textwrap.dedent("""
        import sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import numpy as np
        import adc_testlib as T
        import adcensus_b200 as A
        cases = [T.load_cone() + (64,)]
        for (w, h, D, seed) in ((120, 90, 48, 3), (97, 61, 24, 5)):
            l, r = T.synthetic_pair(w, h, D, seed)
            cases.append((l, r, D))
        for left, right, D in cases:
            h, w, _ = left.shape
            opt = T.default_option(max_disparity=D)
            orc = T.Oracle(w, h, opt)
            eng = A.Engine(w, h, A.ADCensusOption(max_disparity=D))
            orc.begin(left, right); orc.run_to("VOTE")
            eng.debug_run(left, right, "VOTE")
            c = eng.counters()   # [13] = 1 when the adjacency lists were used (they need room in the idle cost volume,
            #                      which small images with long lists do not have -- Cone does)
            if %r: assert c[13] == 0, f"adjacency lists used although ADC_VOTE_ENUM=1: counters {c}"
            elif (w, h) == (450, 375): assert c[13] == 1, f"Cone fell back to enumeration: counters {c}"
            for tap in ("DISP_L", "MISMATCHES", "OCCLUSIONS"):
                g, o = eng.tap(tap), orc.tap(tap)
                assert g.shape == o.shape and g.tobytes() == o.tobytes(), tap
            want = orc.match(left, right)
            assert eng.match(left, right).tobytes() == want.tobytes()
            eng.close()
        print("ok")
    """ % (str(T.REPO), str(T.REPO / "tests"), "ADC_VOTE_ENUM" in env)))
    env = dict(os.environ, **env)
    r = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_pipelined_batches(cone):
    """adc_set_pipelined: three async batch calls in a row without a join in between (different inputs, different output
    buffers, more pairs than one wave), one adc_join at the end; every map must equal the single-pair result."""
    import torch
    left, right = cone
    h, w, _ = left.shape
    eng = _engine(w, h, T.default_option(), wave_pairs=4, lanes=3)
    single = [eng.match(left, right), eng.match(right, left), eng.match(left[:, ::-1].copy(), right[:, ::-1].copy())]
    ins = [(left, right), (right, left), (left[:, ::-1].copy(), right[:, ::-1].copy())]
    n = 10
    st = torch.cuda.current_stream()
    eng.set_pipelined(True)
    outs, keep = [], []
    for l, r in ins:
        dl = torch.from_numpy(np.repeat(l[None], n, 0)).cuda()
        dr = torch.from_numpy(np.repeat(r[None], n, 0)).cuda()
        dd = torch.zeros((n, h, w), dtype=torch.float32, device="cuda")
        keep.append((dl, dr))
        eng.match_batch_device(n, dl.data_ptr(), dr.data_ptr(), dd.data_ptr(), st.cuda_stream)
        outs.append(dd)
    eng.join(st.cuda_stream)
    torch.cuda.synchronize()
    for want, got in zip(single, outs):
        assert (got.cpu().numpy().view(np.uint32) == want.view(np.uint32)[None]).all()
    eng.set_pipelined(False)
    eng.close()


def test_async_refine_lanes(cone):
    """adc_config.async_refine (experimental): a lane's refinement stage on a second stream with its own buffer set while
    the first stream streams the next wave's volumes.  Same bits as the plain schedule, over several waves per lane."""
    import torch
    left, right = cone
    h, w, _ = left.shape
    plain = _engine(w, h, T.default_option(), wave_pairs=4, lanes=2)
    want = [plain.match(left, right), plain.match(right, left)]
    plain.close()
    eng = _engine(w, h, T.default_option(), wave_pairs=4, lanes=2, async_refine=True)
    n = 24
    L = np.stack([left if i % 2 == 0 else right for i in range(n)])
    R = np.stack([right if i % 2 == 0 else left for i in range(n)])
    dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    dd = torch.zeros((n, h, w), dtype=torch.float32, device="cuda")
    for _ in range(2):
        eng.match_batch_device(n, dl.data_ptr(), dr.data_ptr(), dd.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = dd.cpu().numpy()
    for i in range(n):
        assert out[i].tobytes() == want[i % 2].tobytes(), f"pair {i}"
    # the synchronous single-pair entry still works on such an engine (lane buffers, both streams idle)
    assert eng.match(left, right).tobytes() == want[0].tobytes()
    eng.close()
