#!/usr/bin/env python
"""bench.py -- the contract benchmark of the B200 AD-Census engine.

Metric (BASELINE.json): disparity-maps/sec on Cone 450x375x64, batch 256 per GPU (configs[1]).
A "step" = one pass of the whole hot path (ADCensusStereo::Match for every pair) over one batch of
256 stereo pairs.  Inputs are synthetic in the sense of the contract: the bundled Cone pair
replicated 256x; every output therefore has to equal the oracle's map.

  value : whole-job maps/s with the inputs already resident in HBM (adc_match_batch_device)
  e2e   : the same through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the
          timed region, adc_match_batch_pinned_async + synchronise)
  roofline     : the dominant kernel timed in isolation (CUDA events on the engine's stream)
  cpu_baseline : the reference's own CPU implementation (oracle/_ref) on this box's host cores

`--impl reference` times the reference CPU path (oracle/_ref when it was built, else the oracle
port) with one independent instance per host core, on the same workload/metric.

Multi-GPU: independent pairs shard over the ranks (weak scaling, 256 pairs per rank per step);
NCCL carries only the job descriptor broadcast, the timing max and a result checksum.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

WORKLOAD = "cone_450x375_d64_batch256"
PAIRS_PER_STEP = 256
METRIC = "disparity-maps/sec (450x375x64)"


# ------------------------------------------------------------------------------------------------
# reference / CPU arm
def _cpu_worker(args):
    kind, iters = args
    import adc_testlib as T
    left, right = T.load_cone()
    h, w, _ = left.shape
    eng = T.Reference(w, h) if kind == "reference" else T.Oracle(w, h)
    t0 = time.perf_counter()
    eng.time_match(left, right, iters)
    return time.perf_counter() - t0


def cpu_baseline(iters_per_core: int, cores: int | None = None) -> dict:
    """The reference has no threads: the most a box can do with it is one independent single-threaded
    instance per core.  On a many-core host that many instances fight over memory bandwidth, so the
    instance count is scanned (all cores, half, a quarter) and the best aggregate is reported."""
    if cores is None:
        n = os.cpu_count() or 1
        best = None
        tried = {}
        for c in sorted({n, max(1, n // 2), max(1, n // 4)}, reverse=True):
            r = _cpu_baseline_n(iters_per_core, c)
            tried[str(c)] = r["value"]
            if best is None or r["value"] > best["value"]:
                best = r
        best["instances_tried_maps_per_s"] = tried
        return best
    return _cpu_baseline_n(iters_per_core, cores)


def _cpu_baseline_n(iters_per_core: int, cores: int) -> dict:
    import adc_testlib as T
    T.build_oracle()
    kind = "reference" if T.have_ref() else "port"
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        per = pool.map(_cpu_worker, [(kind, iters_per_core)] * cores)
    wall = time.perf_counter() - t0
    busy = max(per)
    return {"value": round(cores * iters_per_core / busy, 4), "unit": "maps/s", "cores": cores, "kind": kind,
            "sample": f"{iters_per_core} Cone 450x375x64 Match calls on each of {cores} independent "
                      f"single-threaded instances ({'oracle/_ref, unmodified reference sources' if kind == 'reference' else 'oracle/adc_oracle.c port'}); "
                      f"slowest instance {busy:.2f}s, wall incl. process start {wall:.2f}s",
            "single_core_s_per_map": round(statistics.median(per) / iters_per_core, 4)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # bounded sample: every "step" = iters_per_core matches on each core.  The instance count (all host threads, half, a
    # quarter -- the single-threaded reference is memory-bound when every hardware thread runs one) is scanned once,
    # untimed; the timed steps then run at the best count, so that the whole arm stays within a few minutes.
    per_step = 1
    scan = cpu_baseline(per_step)
    cores = scan["cores"]
    for _ in range(max(0, args.warmup - 1) and 1 or 0):
        cpu_baseline(per_step, cores)
    t0 = time.perf_counter()
    res = None
    vals = []
    for _ in range(args.steps):
        res = cpu_baseline(per_step, cores)
        vals.append(res["value"])
    res["instances_tried_maps_per_s"] = scan.get("instances_tried_maps_per_s")
    wall = time.perf_counter() - t0
    value = statistics.median(vals)
    res["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "maps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1000 * wall / max(1, args.steps), 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": PAIRS_PER_STEP,
                       "note": "each CPU step is a bounded sample of the workload (one Cone pair per host core)"},
            "cpu_baseline": res,
            "e2e": {"value": value, "unit": "maps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_ev = index, [], threading.Event()

    def run(self):
        while not self._stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop_ev.wait(0.2)

    def stop(self) -> dict:
        self._stop_ev.set()
        self.join(timeout=3)
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import adcensus_b200 as A
    import adc_testlib as T

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # workloads: BASELINE.json configs[1] (the contract's metric) by default; configs[2] / configs[3] on request
    if args.workload == "cone":
        left, right = T.load_cone()
        dmax_w, n, wl_name = 64, PAIRS_PER_STEP, WORKLOAD
        lefts = rights = None
    else:
        w0, h0, dmax_w, n = (1242, 375, 128, 512) if args.workload == "kitti" else (1920, 1080, 192, 64)
        wl_name = f"synthetic_{w0}x{h0}_d{dmax_w}_batch{n}"
        seeds = 16 if args.workload == "kitti" else 8        # distinct synthetic pairs, cycled through the batch
        pairs = [T.synthetic_pair(w0, h0, dmax_w, s + 1) for s in range(seeds)]
        lefts = np.stack([pairs[i % seeds][0] for i in range(n)])
        rights = np.stack([pairs[i % seeds][1] for i in range(n)])
        left, right = pairs[0]
    h, w, _ = left.shape
    # job descriptor from rank 0 (the only data-path collective besides the final reductions)
    desc = torch.tensor([w, h, 0, dmax_w, n], dtype=torch.int32, device=dev)
    if world > 1:
        dist.broadcast(desc, src=0)
    w, h, dmin, dmax, n = [int(v) for v in desc.tolist()]

    opt = A.ADCensusOption(min_disparity=dmin, max_disparity=dmax)
    eng = A.Engine(w, h, opt, device=local, wave_pairs=args.wave_pairs, lanes=args.lanes)
    eng.set_pipelined(not args.no_pipeline)
    N = w * h
    h_left = torch.from_numpy(np.repeat(left[None], n, 0) if lefts is None else lefts).pin_memory()
    h_right = torch.from_numpy(np.repeat(right[None], n, 0) if rights is None else rights).pin_memory()
    h_disp = torch.empty((n, h, w), dtype=torch.float32).pin_memory()
    d_left, d_right = h_left.to(dev), h_right.to(dev)
    d_disp = torch.empty((n, h, w), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count
        e0.record(st)
        for _ in range(steps):
            fn()
        eng.join(st.cuda_stream)        # pipelined engine: the K steps flow into each other, the join is inside the timed region
        e1.record(st)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        launches = torch.tensor([eng.launch_count - l0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(launches, op=dist.ReduceOp.SUM)
        barrier()
        return float(ms.item()), int(launches.item())

    dev_step = lambda: eng.match_batch_device(n, d_left.data_ptr(), d_right.data_ptr(), d_disp.data_ptr(), st.cuda_stream)
    e2e_step = lambda: eng.match_batch_pinned_async(n, h_left.data_ptr(), h_right.data_ptr(), h_disp.data_ptr(), st.cuda_stream)

    for _ in range(max(3, args.warmup)):
        dev_step()
    eng.join(st.cuda_stream)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev, launches = timed(dev_step, args.steps)
    clocks = sampler.stop() if sampler else None
    for _ in range(2):
        e2e_step()
    eng.join(st.cuda_stream)
    ms_e2e, _ = timed(e2e_step, args.steps)

    # correctness guard inside the bench: every map of the batch must equal the single-pair result
    ref_map = eng.match(left, right)
    torch.cuda.synchronize()
    if lefts is None:
        ok = bool((h_disp.numpy().view(np.uint32) == ref_map.view(np.uint32)[None]).all())
        okd = bool((d_disp.cpu().numpy().view(np.uint32) == ref_map.view(np.uint32)[None]).all())
    else:   # cycled distinct pairs: every copy of pair 0 must equal its single-pair map, and both paths must agree
        idx = np.arange(0, n, seeds)
        ok = bool((h_disp.numpy()[idx].view(np.uint32) == ref_map.view(np.uint32)[None]).all())
        okd = bool(np.array_equal(d_disp.cpu().numpy().view(np.uint32), h_disp.numpy().view(np.uint32)))
    flag = torch.tensor([int(ok and okd)], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)

    line = None
    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        kern = {}
        for name in ("cost_volume", "arm_sum_h", "arm_sum_v_div", "scanline_x", "scanline_y", "wta"):
            kms, kbytes = eng.profile_kernel(name, reps=5)
            kern[name] = {"ms_per_launch": round(kms, 4), "algorithmic_bytes": kbytes,
                          "achieved_gbs": round(kbytes / (kms * 1e-3) / 1e9, 1),
                          "frac": round(kbytes / (kms * 1e-3) / 1e9 / hbm_peak, 4), "pairs_per_launch": eng.wave_pairs}
        # dominant kernel of a step = the one with the largest share of device time:
        # 8 arm-sum launches and 4 scanline launches per wave
        share = {"arm_sum": 4 * (kern["arm_sum_h"]["ms_per_launch"] + kern["arm_sum_v_div"]["ms_per_launch"]),
                 "scanline": 2 * (kern["scanline_x"]["ms_per_launch"] + kern["scanline_y"]["ms_per_launch"])}
        dom = "scanline_x" if share["scanline"] >= share["arm_sum"] else "arm_sum_v_div"
        # measured DRAM bytes of that kernel (ncu --set full capture summarised in profiles/; per pair there, per launch here)
        traffic = None
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists():
            per_pair = json.loads(tf.read_text()).get(dom, {}).get("dram_bytes_per_pair")
            if per_pair:
                traffic = round(per_pair * eng.wave_pairs)
        roof = {"kernel": dom, "bound": "hbm", "achieved": kern[dom]["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                "frac": kern[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "note": "algorithmic bytes per launch = (2V + 6N) per pair x pairs per launch, V = 4*H*W*D (SURVEY 8d); "
                        "traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture "
                        "(profiles/r1_ncu_full_v32_summary.csv); kernel timed alone with CUDA events on the engine's stream"}
        total_maps = world * n * args.steps
        value = total_maps / (ms_dev * 1e-3)
        e2e_v = total_maps / (ms_e2e * 1e-3)
        b_map = 18.0 * 4.0 * N * (dmax - dmin)
        cpu = cpu_baseline(1) if (world == 1 and not args.no_cpu and args.workload == "cone") else None
        metric = METRIC if args.workload == "cone" else f"disparity-maps/sec ({w}x{h}x{dmax - dmin})"
        line = {"metric": metric, "value": round(value, 2), "unit": "maps/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": round(ms_dev / args.steps, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl_name, "pairs_per_step_per_gpu": n, "width": w, "height": h, "disparities": dmax - dmin,
                           "wave_pairs": eng.wave_pairs, "lanes": eng.lanes,
                           "steps_pipelined": not args.no_pipeline,
                           "l2_policy": "no flush needed: each step streams the whole batch of images and several GB of cost volumes per wave, far beyond the 126 MB L2",
                           "parallelism": f"dp{world} (independent pairs, no data-path collective)"},
                "e2e": {"value": round(e2e_v, 2), "unit": "maps/s", "h2d_bytes_per_step": n * 2 * N * 3,
                        "d2h_bytes_per_step": n * N * 4, "ms_per_step": round(ms_e2e / args.steps, 3)},
                "gpu_launches": launches, "clocks": clocks, "outputs_bit_identical": bool(flag.item()),
                "pipeline_hbm": {"algorithmic_bytes_per_map": b_map, "achieved_gbs": round(value * b_map / 1e9 / world, 1),
                                 "frac": round(value * b_map / 1e9 / world / hbm_peak, 4)},
                "roofline": roof, "kernels": kern}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--wave-pairs", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="join the stream after every step (adc_set_pipelined off): each step then drains the engine")
    ap.add_argument("--workload", default="cone", choices=["cone", "kitti", "1080p"],
                    help="cone = BASELINE configs[1] (the contract metric); kitti / 1080p = configs[2] / configs[3] (extra lines)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
