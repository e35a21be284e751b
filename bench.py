#!/usr/bin/env python
"""
bench.py -- octvr stitch hot path on B200: equirect output Mpix/s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3] [--impl ours|reference]

A "step" is one stitched output frame: N camera frames (synthetic, device-resident ring of 8 distinct
frames per camera) -> one equirectangular 4:2:0 frame through liboctvr_b200.so.  N=1 workload = C2
(6 x 2704x1520 -> 4096x2048, gain + feather) -- the configuration BASELINE.json's target is quoted on.
With --gpus N > 1 (torchrun, one rank per GPU) every rank stitches its own stream (frame sharding,
SURVEY.md 8e: no data-path collective), value = frames of all ranks / max-over-ranks time.
--impl reference times the CPU oracle port of the reference's CPU path on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (rig, blend, gain, description)
    "c1": ("rig2", -1, True, "C1 2x1920x1080 fisheye -> 2048x1024 equirect, gain + feather(1)"),
    "c2": ("rig6", -1, True, "C2 6x2704x1520 -> 4096x2048 equirect, gain + feather(1)"),
    "c3": ("rig6", 64, True, "C3 6x2704x1520 -> 4096x2048 equirect, gain + 5-band multiband"),
    "c4": (("rig8L", "rig8R"), 64, True, "C4 8x3840x2160 fisheye -> 7680x3840 stereo top-bottom, gain + 5-band multiband"),
    "c5": ("rig6", -1, True, "C5 16 independent streams of the C2 rig (6x2704x1520 -> 4096x2048, gain + feather(1)), frame-sharded"),
    "fast": ("rig6", -5, False, "FastMapper (NV12, u8 feather weights) on the C2 rig: 6x2704x1520 NV12 -> 4096x2048 NV12"),
    "c2ng": ("rig6", -1, False, "C2 rig without gain compensation: 6x2704x1520 -> 4096x2048 equirect, feather(1) (diagnostic)"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md).  The timed region of this
    bench is tens of milliseconds, too short for `nvidia-smi -lms`, so NVML is polled directly (every ~1 ms) from a
    thread; nvidia-smi is the fallback."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.sm, self.mask, self.max_mhz, self.err = [], 0, None, None
        self._stop = threading.Event()
        self.t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception as ex:          # noqa: BLE001
            self.err = "nvml: %s" % ex

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:        # noqa: BLE001  (older bindings)
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception as ex:      # noqa: BLE001
                self.err = "nvml: %s" % ex
                return
            time.sleep(0.001)

    def stop(self):
        self._stop.set()
        if self.t:
            self.t.join(timeout=2)
        if not self.sm:
            return self._smi_once()
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for bit, n in self.REASONS.items() if self.mask & bit), "samples": len(self.sm), "source": "nvml"}

    def _smi_once(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            r = [c.strip() for c in out.splitlines()[0].split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(r[0]), "sm_max_mhz": float(r[1]), "reasons": [names[k] for k in range(4) if r[2 + k].lower().startswith("active")],
                    "samples": 1, "source": "nvidia-smi after the timed region (%s)" % self.err}
        except Exception as ex:          # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable: %s / %s" % (self.err, ex)], "samples": 0}


def alg_bytes(in_sizes, out_size, pairs, roi_area, blend):
    """B_alg = I + T + O (+ M for multiband)."""
    I = sum(w * h * 3 // 2 for w, h in in_sizes)
    O_ = out_size[0] * out_size[1] * 3 // 2
    if blend > 0:
        M = 2 * (4 / 3) * roi_area * 4 + 2 * (4 / 3) * out_size[0] * out_size[1] * 6 + (4 / 3) * roi_area * 4
        return I + 8 * pairs + O_ + M
    return I + 12 * pairs + O_


def blend_kernel_name():
    return {"fused": "k_stitch_fused", "direct": "k_blend", "staged": "k_blend_staged"}.get(os.environ.get("OCTVR_BLEND", ""), "k_blend_ring")


def ncu_traffic(workload, blend, kernel=None):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch.  NOT measured in this run (ncu is
    never active while timing): read from the committed ncu --set full capture of this workload
    (profiles/r02_traffic.json, else r01_traffic.json; written from the .ncu-rep)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
            v = t.get("%s:%s" % (workload, kernel or (blend_kernel_name() if blend <= 0 else "multiband")))
            if v is not None:
                return v
        except Exception:                    # noqa: BLE001
            pass
    return None


_PG = {"up": False}


def pg_init(local):
    """One NCCL process group per process, shared by the headline workload and the sub-benchmarks.  NCCL's stream is
    created high-priority: the row-band exchange (C4) runs next to the stitch kernels and must not queue behind them."""
    import torch
    import torch.distributed as dist
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not _PG["up"]:
        hp = os.environ.get("OCTVR_NCCL_HP", "1") != "0"
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=hp)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
        _PG["up"] = True


def pg_done():
    import torch.distributed as dist
    if _PG["up"]:
        dist.barrier()
        dist.destroy_process_group()
        _PG["up"] = False


def make_template(vr, cfg, width, device):
    """Tables come from the product's own GPU map generation (octvr_template_build_json)."""
    return vr.MapperTemplate.from_json(cfg, width, -1, use_roi=True, with_seam_masks=True, device=device)


def measure_single(args, workload, steps, warmup, with_e2e):
    """One stream per rank of a single-template workload (c1, c2, c3): device-resident timing + optional e2e.  Returns the
    bench line as a dict on rank 0 (None elsewhere).  Collective: every rank calls it."""
    import torch
    import torch.distributed as dist
    import octvr_b200 as vr
    import util

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    pg_init(local)
    torch.cuda.set_device(local)
    rig, blend, gain, desc = WORKLOADS[workload]
    cfg, width, in_size = util.named_rig(rig)
    n = len(cfg["inputs"])
    iw, ih = in_size
    t0 = time.time()
    tmpl = make_template(vr, cfg, width, local)
    t_tmpl = time.time() - t0
    t0 = time.time()
    m = vr.Mapper(tmpl, [in_size] * n, blend=blend, enable_gain_compensator=gain, device=local)
    t_map = time.time() - t0
    W, H = tmpl.out_size
    st = m.stats()

    RING = 8
    ring = []
    for k in range(RING):     # distinct noise frames, Mapper's packed layout, resident in HBM
        fr = []
        for c in range(n):
            y, u, v = util.i420_planes(util.noise_frame(c, iw, ih, seed=1234 + 7919 * k + 104729 * rank), iw, ih)
            fr.append(torch.from_numpy(np.concatenate([y, np.concatenate([u, v], 1)], 0)).cuda())
        ring.append(fr)
    out = torch.zeros((H * 3 // 2, W), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()

    def step(k):
        m.stitch_packed(ring[k % RING], out, stream=stream)

    for k in range(warmup):
        step(k)
    torch.cuda.synchronize()

    # per-stage split (separate, untimed pass)
    m.set_profiling(True)
    stage = {"convert": [], "gain": [], "blend": []}
    for k in range(10):
        step(k)
        for s in stage:
            stage[s].append(m.stage_ms(s))
    m.set_profiling(False)
    # (the sampler starts BEFORE the barrier, see run_stereo)
    sampler = ClockSampler(local) if rank == 0 else None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(steps):
        step(k)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    ms = vr.sharding.max_over_ranks(ms, device="cuda")            # timing rule: max over ranks
    frames_all = sum(vr.sharding.gather_frame_counts(steps, device="cuda"))

    # end to end through the host-facing API (AsyncMultiMapper: host planes in, host planes out)
    e2e = None
    if with_e2e:
        try:
            e2e = run_e2e(vr, util, tmpl, cfg, in_size, blend, gain, local, min(steps, 60), world)
        except vr.OctvrError as ex:
            e2e = {"value": None, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(ex)}
    del m, ring, out
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    ms_per_step = ms / steps
    frames = frames_all
    mpix = W * H * frames / (ms * 1e-3) / 1e6
    B = alg_bytes([in_size] * n, (W, H), st["pairs"], st["roi_area"], blend)
    peak, how = peaks()
    # dominant kernel = the fused blend kernel; its algorithmic bytes = tables + output (the convert
    # kernel owns the input bytes), duration from CUDA events on the launch stream
    blend_ms = statistics.median(stage["blend"])
    if blend > 0:    # multiband stage (k_mb_warp .. k_mb_final): everything of B_alg except the input frames
        blend_bytes = B - n * iw * ih * 3 // 2
    else:
        blend_bytes = 12 * st["pairs"] + W * H * 3 // 2
    ach = blend_bytes / (blend_ms * 1e-3) / 1e9
    line = {
        "metric": "equirect output Mpix/s", "value": round(mpix, 1), "unit": "Mpix/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": round(ms_per_step, 5), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "frames_per_s": round(frames / (ms * 1e-3), 1),
        "config": {"workload": desc, "inputs": "%dx%dx%d I420 (octvr packed layout), ring of %d distinct noise frames per camera resident in HBM" % (n, iw, ih, RING),
                   "output": "%dx%d 4:2:0" % (W, H), "l2": "working set per step (%.0f MB tables + %.0f MB frames) exceeds the 126 MB L2; no flush" % (st["table_bytes"] / 1e6, n * iw * ih * 1.5 / 1e6),
                   "pairs_P": st["pairs"], "roi_area": st["roi_area"], "sharding": "one independent stream per GPU (frame sharding, no collective)" if world > 1 else "single GPU",
                   "template_build_s": round(t_tmpl, 2), "mapper_build_s": round(t_map, 2)},
        "alg_bytes_per_frame": int(B), "table_bytes_per_frame": st["table_bytes"],
        "frac_of_hbm_roofline": {"whole_step_vs_measured_%.0f" % peak: round(B / (ms_per_step * 1e-3) / 1e9 / peak, 4),
                                 "whole_step_vs_8000": round(B / (ms_per_step * 1e-3) / 1e9 / 8000.0, 4)},
        "stage_ms": {k: round(statistics.median(v), 5) for k, v in stage.items()},
        "roofline": {"bound": "hbm", "kernel": blend_kernel_name() if blend <= 0 else "multiband stage (k_mb_warp+k_mb_down+k_mb_band+k_mb_collapse+k_mb_final)", "achieved": round(ach, 1), "peak": peak, "peak_source": how, "unit": "GB/s",
                     "frac": round(ach / peak, 4), "traffic": ncu_traffic(workload, blend),
                     "traffic_source": "committed ncu --set full capture under profiles/ (not measured in this run)",
                     "alg_bytes_per_launch": int(blend_bytes), "ms_per_launch": round(blend_ms, 5)},
        "gpu_launches": st["launches_per_stitch"] * steps,
        "clocks": clocks,
    }
    if with_e2e:
        line["e2e"] = e2e
    return line


def sub_summary(line, keys=("value", "unit", "ms_per_step", "frames_per_s", "n_gpus", "steps", "warmup", "scaling", "frac_of_hbm_roofline", "roofline",
                            "stage_ms", "stage_ms_rank0", "stitch_only_ms_max_over_ranks", "host_enqueue_ms_per_step_rank0", "assembled_frame_equals_single_gpu_result",
                            "exchange_alone_ms", "gpu_launches", "ms_per_frame_per_gpu", "clocks")):
    if line is None:
        return None
    d = {k: line[k] for k in keys if k in line}
    d["workload"] = line.get("config", {}).get("workload")
    for k in ("sharding", "streams_per_rank", "template_build_s", "mapper_build_s"):
        if k in line.get("config", {}):
            d[k] = line["config"][k]
    return d


def run_ours(args):
    """Headline line = the named workload (default C2).  With the default workload the same run also measures the other
    BASELINE configurations as sub-objects of the line, so the driver's BENCH / SCALE files carry them: `c3` (5-band
    multiband, one stream per GPU), `c5` (16 streams frame-sharded over the GPUs) and `c4_rowband` (ONE 7680x3840 stereo
    stream split by (eye, row band) over all GPUs, strong scaling, assembled frame verified against the single-GPU frame)."""
    import argparse as _ap
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = measure_single(args, args.workload, args.steps, args.warmup, with_e2e=True)
    subs = {}
    if args.workload == "c2" and not args.no_sub:
        def guarded(name, fn):
            try:
                subs[name] = sub_summary(fn())
            except Exception as ex:          # noqa: BLE001  (a failed sub-benchmark must not take the headline line with it)
                subs[name] = {"error": "%s: %s" % (type(ex).__name__, ex)}
        guarded("c3", lambda: measure_single(args, "c3", max(10, min(args.steps, 100)), max(3, min(args.warmup, 10)), with_e2e=False))
        a5 = _ap.Namespace(**vars(args)); a5.workload = "c5"; a5.steps = max(3, min(args.steps, 8)); a5.warmup = 3
        guarded("c5", lambda: run_streams(a5, sub=True))
        af = _ap.Namespace(**vars(args)); af.workload = "fast"; af.steps = max(10, min(args.steps, 100))
        guarded("fast_nv12", lambda: run_fast(af, sub=True))
        a4 = _ap.Namespace(**vars(args)); a4.workload = "c4"; a4.steps = max(5, min(args.steps, 30)); a4.warmup = 3; a4.verify = world > 1
        guarded("c4_rowband", lambda: run_stereo(a4, sub=True))
    pg_done()
    if rank != 0:
        return
    line.update(subs)
    if not args.no_cpu and world == 1:      # the CPU baseline is timed on rank 0 at N = 1 only
        line["cpu_baseline"] = cpu_baseline(args.workload, budget_s=20.0)
    print(json.dumps(line))


def run_fast(args, sub=False):
    """vr::FastMapper (mapper_fast.cpp) on the C2 rig with full-frame tables (the reference's FastMapper takes no ROI): NV12
    frames resident in HBM, one launch of k_fast_nv12 per frame.  Frame-sharded like the headline workload (each rank its own
    stream, no data-path collective).  Algorithmic bytes per frame: I + 8 * (P_luma + P_chroma) + O (8-byte table entries for
    the contributing (pixel, camera) pairs of the luma and the chroma tables)."""
    import torch
    import torch.distributed as dist
    import octvr_b200 as vr
    import util
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    pg_init(local)
    torch.cuda.set_device(local)
    rig, _, _, desc = WORKLOADS["fast"]
    cfg, width, in_size = util.named_rig(rig)
    n = len(cfg["inputs"])
    iw, ih = in_size
    t0 = time.time()
    tmpl = vr.MapperTemplate.from_json(cfg, width, -1, use_roi=False, with_seam_masks=False, device=local)
    t_tmpl = time.time() - t0
    t0 = time.time()
    fm = vr.FastMapper(tmpl, [in_size] * n, device=local)
    t_map = time.time() - t0
    W, H = tmpl.out_size
    RING = 8
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    frames = [[torch.randint(0, 256, (ih + ih // 2, iw), dtype=torch.uint8, device="cuda", generator=gen) for c in range(n)] for k in range(RING)]
    out = torch.zeros((H + H // 2, W), dtype=torch.uint8, device="cuda")
    for k in range(max(3, args.warmup)):
        fm.stitch_nv12(frames[k % RING], out)
    sampler = ClockSampler(local) if rank == 0 else None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        fm.stitch_nv12(frames[k % RING], out)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms = vr.sharding.max_over_ranks(e0.elapsed_time(e1), device="cuda")
    info = fm.info()
    del fm, frames, out
    torch.cuda.empty_cache()
    if not sub:
        pg_done()
    if rank != 0:
        return None
    peak, how = peaks()
    I, O_ = n * iw * (ih + ih // 2), W * (H + H // 2)
    B = I + 8 * (info["pairs_luma"] + info["pairs_chroma"]) + O_
    ms_per_step = ms / args.steps
    ach = B / (ms_per_step * 1e-3) / 1e9
    fast_kernel = "k_fast_nv12" if os.environ.get("OCTVR_FAST", "") == "direct" else "k_fast_staged"
    return {
        "metric": "equirect output Mpix/s", "value": round(world * W * H / (ms_per_step * 1e-3) / 1e6, 1), "unit": "Mpix/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "frames_per_s": round(world * args.steps / (ms * 1e-3), 1),
        "config": {"workload": desc, "inputs": "%dx%dx%d NV12, ring of %d noise frames per camera resident in HBM" % (n, iw, ih, RING),
                   "output": "%dx%d NV12" % (W, H), "l2": "tables (%.0f MB) + frames exceed the 126 MB L2; no flush" % (info["table_bytes"] / 1e6),
                   "pairs_luma": info["pairs_luma"], "pairs_chroma": info["pairs_chroma"], "sharding": "one stream per GPU, no data-path collective",
                   "template_build_s": round(t_tmpl, 2), "mapper_build_s": round(t_map, 2)},
        "alg_bytes_per_frame": int(B),
        "roofline": {"bound": "hbm", "kernel": fast_kernel, "achieved": round(ach, 1), "peak": peak, "peak_source": how, "unit": "GB/s",
                     "frac": round(ach / peak, 4), "traffic": ncu_traffic("fast", 0, fast_kernel),
                     "traffic_source": "committed ncu --set full capture under profiles/ (not measured in this run)",
                     "alg_bytes_per_launch": int(B), "ms_per_launch": round(ms_per_step, 5)},
        "gpu_launches": args.steps, "clocks": clocks,
    }


def run_rowband(args):
    """One stream over all ranks: rank 0 ingests the frames, NCCL broadcasts them, rank r stitches output rows
    row_bands(H, world, 32)[r], the bands are collected on rank 0.  Everything is inside the timed region."""
    import torch
    import torch.distributed as dist
    import octvr_b200 as vr
    import util
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    pg_init(local)
    torch.cuda.set_device(local)
    rig, blend, gain, desc = WORKLOADS[args.workload]
    cfg, width, in_size = util.named_rig(rig)
    n = len(cfg["inputs"])
    iw, ih = in_size
    tmpl = make_template(vr, cfg, width, local)
    W, H = tmpl.out_size
    rb = vr.sharding.RowBandStitcher(vr, tmpl, [in_size] * n, blend, gain, local)
    RING = 4
    ring, flats = [], []
    for k in range(RING):
        flat, views = vr.sharding.alloc_frame_set([in_size] * n, "cuda")
        if rank == 0:
            for c in range(n):
                y, u, v = util.i420_planes(util.noise_frame(c, iw, ih, seed=1234 + 7919 * k), iw, ih)
                views[c].copy_(torch.from_numpy(np.concatenate([y, np.concatenate([u, v], 1)], 0)))
        ring.append(views)
        flats.append(flat)
    out = torch.zeros((H * 3 // 2, W), dtype=torch.uint8, device="cuda")
    pipe = vr.sharding.FramePipeline(rb, src=0)

    def step(k, last=False):
        pipe.step(flats[k % RING], ring[k % RING], out, next_flat=None if last else flats[(k + 1) % RING])

    for k in range(args.warmup):
        step(k, last=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        step(k, last=(k == args.steps - 1))
    e1.record()
    torch.cuda.synchronize()
    ms = vr.sharding.max_over_ranks(e0.elapsed_time(e1), device="cuda")
    pg_done()
    if rank != 0:
        return
    print(json.dumps({
        "metric": "equirect output Mpix/s", "value": round(W * H * args.steps / (ms * 1e-3) / 1e6, 1), "unit": "Mpix/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 5), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "frames_per_s": round(args.steps / (ms * 1e-3), 1),
        "config": {"workload": desc, "sharding": "row bands of one stream: one NCCL broadcast of the %d input frames (%.1f MB) per step, issued one step ahead, + band "
                   "stitch + bands sent to rank 0 (NCCL send/recv)" % (n, n * iw * ih * 1.5 / 1e6), "bands": rb.bands}}))


def run_streams(args, sub=False):
    """BASELINE config C5: 16 independent video streams of the C2 rig, sharded over the ranks (stream s -> rank s mod N, no
    data-path collective).  A rank serves its streams round-robin on `--concurrency` CUDA streams, one Mapper each (a
    Mapper's per-frame buffers belong to one frame at a time), so the latency-bound gain chain of one frame runs under the
    blend of another.  A step = one frame of every stream of the job; value = frames of all ranks / max-over-ranks time."""
    import torch
    import torch.distributed as dist
    import octvr_b200 as vr
    import util
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    pg_init(local)
    torch.cuda.set_device(local)
    rig, blend, gain, desc = WORKLOADS[args.workload]
    cfg, width, in_size = util.named_rig(rig)
    n = len(cfg["inputs"])
    iw, ih = in_size
    NS = 16
    mine = vr.sharding.streams_of_rank(NS, rank, world)
    conc = max(1, min(args.concurrency, len(mine)))
    tmpl = make_template(vr, cfg, width, local)
    W, H = tmpl.out_size
    mappers = [vr.Mapper(tmpl, [in_size] * n, blend=blend, enable_gain_compensator=gain, device=local) for _ in range(conc)]
    cstreams = [torch.cuda.Stream(device=local) for _ in range(conc)]
    st = mappers[0].stats()
    RING = 1 if sub else 2
    frames, outs = {}, {}
    for s_ in mine:            # every video stream has its own frames and its own output
        frames[s_] = []
        for k in range(RING):
            fr = []
            for c in range(n):
                y, u, v = util.i420_planes(util.noise_frame(c, iw, ih, seed=1234 + 7919 * k + 104729 * s_), iw, ih)
                fr.append(torch.from_numpy(np.concatenate([y, np.concatenate([u, v], 1)], 0)).cuda())
            frames[s_].append(fr)
        outs[s_] = torch.zeros((H * 3 // 2, W), dtype=torch.uint8, device="cuda")
    main = torch.cuda.current_stream()

    def step(k):
        for i, s_ in enumerate(mine):
            q = (k * len(mine) + i) % conc
            mappers[q].stitch_packed(frames[s_][k % RING], outs[s_], stream=cstreams[q])

    def fork():
        ev = torch.cuda.Event()
        ev.record(main)
        for cs in cstreams:
            cs.wait_event(ev)

    def join():
        for cs in cstreams:
            ev = torch.cuda.Event()
            ev.record(cs)
            main.wait_event(ev)

    fork()
    for k in range(args.warmup):
        step(k)
    join()
    torch.cuda.synchronize()
    # dominant kernel alone (one mapper, serial): duration for the roofline object
    mappers[0].set_profiling(True)
    stage = {"convert": [], "gain": [], "blend": []}
    for k in range(10):
        mappers[0].stitch_packed(frames[mine[0]][k % RING], outs[mine[0]], stream=main)
        for s_ in stage:
            stage[s_].append(mappers[0].stage_ms(s_))
    mappers[0].set_profiling(False)
    # (the sampler starts BEFORE the barrier: its start-up on rank 0 must not fall between the barrier and the first timed step,
    # where the other ranks of a collective workload would wait for it inside their timed region)
    sampler = ClockSampler(local) if rank == 0 else None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    fork()
    for k in range(args.steps):
        step(k)
    join()
    e1.record(main)
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms = vr.sharding.max_over_ranks(e0.elapsed_time(e1), device="cuda")
    frames_all = sum(vr.sharding.gather_frame_counts(args.steps * len(mine), device="cuda"))
    e2e = None if sub else run_e2e(vr, util, tmpl, cfg, in_size, blend, gain, local, min(args.steps * len(mine), 60), world)
    del mappers, frames, outs
    torch.cuda.empty_cache()
    if not sub:
        pg_done()
    if rank != 0:
        return None
    B = alg_bytes([in_size] * n, (W, H), st["pairs"], st["roi_area"], blend)
    peak, how = peaks()
    per_frame_ms = ms / (args.steps * len(mine))
    blend_ms = statistics.median(stage["blend"])
    blend_bytes = 12 * st["pairs"] + W * H * 3 // 2
    ach = blend_bytes / (blend_ms * 1e-3) / 1e9
    line = {
        "metric": "equirect output Mpix/s", "value": round(W * H * frames_all / (ms * 1e-3) / 1e6, 1), "unit": "Mpix/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 5), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "frames_per_s": round(frames_all / (ms * 1e-3), 1),
        "config": {"workload": desc, "streams": NS, "streams_per_rank": len(mine), "cuda_streams_per_rank": conc,
                   "inputs": "per stream a ring of %d distinct noise frame sets resident in HBM" % RING,
                   "l2": "working set per step (tables + %d streams' frames) exceeds the 126 MB L2; no flush" % len(mine),
                   "sharding": "stream s -> rank s mod N, no data-path collective; a step = one frame of each of the 16 streams",
                   "pairs_P": st["pairs"], "roi_area": st["roi_area"]},
        "alg_bytes_per_frame": int(B), "ms_per_frame_per_gpu": round(per_frame_ms, 5),
        "frac_of_hbm_roofline": {"per_gpu_vs_measured_%.0f" % peak: round(B / (per_frame_ms * 1e-3) / 1e9 / peak, 4)},
        "stage_ms_serial": {k: round(statistics.median(v), 5) for k, v in stage.items()},
        "roofline": {"bound": "hbm", "kernel": blend_kernel_name(), "achieved": round(ach, 1), "peak": peak, "peak_source": how, "unit": "GB/s",
                     "frac": round(ach / peak, 4), "traffic": ncu_traffic("c2", blend), "alg_bytes_per_launch": int(blend_bytes),
                     "ms_per_launch": round(blend_ms, 5), "note": "kernel timed alone on one CUDA stream"},
        "gpu_launches": st["launches_per_stitch"] * args.steps * len(mine), "clocks": clocks, "e2e": e2e}
    if sub:
        return line
    print(json.dumps(line))


def run_stereo(args, sub=False):
    """BASELINE config C4: 8 x 3840x2160 fisheye -> 7680x3840 stereo top-bottom (two 7680x1920 eye templates, 5-band multiband),
    ONE stream over all ranks split by (eye, row band) -- sharding.StereoRowBandStitcher.  N > 1: rank 0 ingests, NCCL
    broadcasts the inputs, the bands are collected on rank 0; everything inside the timed region (strong scaling)."""
    import torch
    import torch.distributed as dist
    import octvr_b200 as vr
    import util
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    # (the exchange runs on NCCL's own stream next to the stitch kernels: pg_init gives it priority, or its few CTAs queue
    # behind the stitch grids and the "overlapped" broadcast only starts when the stitch has drained)
    pg_init(local)
    torch.cuda.set_device(local)
    rigs, blend, gain, desc = WORKLOADS[args.workload]
    eye_h = 1920
    tmpls, in_size = [], None
    t0 = time.time()
    todo = sorted({e for e, _, _ in vr.sharding.stereo_assignment(world)[rank]})
    if args.verify and rank == 0:
        todo = [0, 1]
    from concurrent.futures import ThreadPoolExecutor
    cfgs = [util.named_rig(rig) for rig in rigs]
    n, in_size = len(cfgs[0][0]["inputs"]), cfgs[0][2]
    # only the eye(s) this rank stitches need tables; the other slot reuses the same template object (never stitched).
    # The two eyes are independent: built side by side (the C ABI calls release the GIL)
    with ThreadPoolExecutor(max_workers=2) as ex:
        futs = [ex.submit(vr.MapperTemplate.from_json, cfg, width, eye_h, True, True, local) if e in todo else None
                for e, (cfg, width, _) in enumerate(cfgs)]
        tmpls = [f.result() if f is not None else None for f in futs]
    for e in range(2):
        if tmpls[e] is None:
            tmpls[e] = tmpls[todo[0]]
    t_tmpl = time.time() - t0
    iw, ih = in_size
    t0 = time.time()
    peer = world > 1 and os.environ.get("OCTVR_C4_COLLECT", "peer") == "peer"
    st = vr.sharding.StereoRowBandStitcher(vr, tmpls, [in_size] * n, blend, gain, local, split="auto" if peer or world == 1 else "rows")
    t_map = time.time() - t0
    W, He = st.eye_w, st.eye_h
    H = 2 * He
    RING = 4
    ring, flats = [], []
    for k in range(RING):      # every time step's frames in one contiguous buffer: a single broadcast per step
        flat, views = vr.sharding.alloc_frame_set([in_size] * n, "cuda")
        if rank == 0:
            for c in range(n):
                y, u, v = util.i420_planes(util.noise_frame(c, iw, ih, seed=1234 + 7919 * k), iw, ih)
                views[c].copy_(torch.from_numpy(np.concatenate([y, np.concatenate([u, v], 1)], 0)))
        ring.append(views)
        flats.append(flat)
    # N > 1: the two output frames live on rank 0 and are mapped into every rank (CUDA IPC over NVLink): each rank's blend
    # kernels store their band straight into rank 0's frame, one 4-byte all-reduce per step says "frame complete".
    # OCTVR_C4_COLLECT=nccl: private output buffers, bands sent to rank 0 by NCCL send / recv (the round-1 path).
    pfs = [vr.sharding.PeerFrame(H * 3 // 2, W, local, owner=0) for _ in range(2)] if peer else []
    outs = [p.tensor for p in pfs] if peer else [torch.zeros((H * 3 // 2, W), dtype=torch.uint8, device="cuda") for _ in range(2)]
    out = outs[0]
    pipe = vr.sharding.FramePipeline(st, src=0, defer_collect=os.environ.get("OCTVR_DEFER_COLLECT", "0") != "0", peer=peer)

    # N > 1: only the source columns some mapper reads travel (the C4 rig's fisheye circles use 56 % of each 3840-wide frame):
    # per step rank 0 crops the full frames of the NEXT time step into a compact frame set (one launch, side stream) and that
    # is what is broadcast; the mappers are told their frames start at the window's first column.  OCTVR_C4_CROP=0: full frames.
    crop = world > 1 and os.environ.get("OCTVR_C4_CROP", "1") != "0"
    bcast_bytes = sum(f.numel() for f in flats[:1])
    if crop:
        cols = st.source_cols()
        lo = torch.tensor([c[0] for c in cols], dtype=torch.int32, device="cuda")
        hi = torch.tensor([c[1] for c in cols], dtype=torch.int32, device="cuda")
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        windows = [(int(a) // 32 * 32, (int(b) + 31) // 32 * 32 - int(a) // 32 * 32) for a, b in zip(lo.tolist(), hi.tolist())]
        st.set_input_windows(windows)
        full_ring, full_flats = ring, flats
        ring, flats = [], []
        for k in range(RING):
            flat, views = vr.sharding.alloc_frame_set([(cw, in_size[1]) for _, cw in windows], "cuda")
            if rank == 0:
                vr.sharding.crop_packed_frames(full_ring[k], views, [in_size] * n, windows)
            dist.broadcast(flat, 0)                    # every rank starts with valid frames in every slot (warm-up, stage timing)
            ring.append(views)
            flats.append(flat)
        by_ptr = {flats[k].data_ptr(): k for k in range(RING)}
        pipe.prepare = lambda flat: vr.sharding.crop_packed_frames(full_ring[by_ptr[flat.data_ptr()]], ring[by_ptr[flat.data_ptr()]], [in_size] * n, windows)
        bcast_bytes = flats[0].numel()
    nobcast = os.environ.get("OCTVR_C4_NOBCAST", "0") != "0"      # diagnostic: frames taken as already resident on every rank
    if nobcast and world > 1:
        for fl in flats:
            dist.broadcast(fl, 0)
        vr.sharding.broadcast_frames = lambda *a, **k: []

    def step(k, last=False):
        if world > 1:      # broadcast of step k + 1 runs under the stitch of step k (OCTVR_DEFER_COLLECT=1: the band collection of step k - 1 too)
            pipe.step(flats[k % RING], ring[k % RING], outs[k % 2], next_flat=None if last else flats[(k + 1) % RING])
            if last:
                pipe.flush()
        else:
            st.stitch_local(ring[k % RING], out)

    for k in range(args.warmup):
        step(k, last=True)
    torch.cuda.synchronize()
    # per-stage split and stitch-only time of this rank (separate, untimed pass)
    stage = {"convert": [], "gain": [], "blend": [], "total": []}
    for _, _, m in st.jobs:
        m.set_profiling(True)
    for k in range(5):
        st.stitch_local(ring[k % RING], out)
        for s_ in stage:
            stage[s_].append(sum(m.stage_ms(s_) for _, _, m in st.jobs))
    for _, _, m in st.jobs:
        m.set_profiling(False)
    # (the sampler starts BEFORE the barrier: its start-up on rank 0 must not fall between the barrier and the first timed step,
    # where the other ranks of a collective workload would wait for it inside their timed region)
    sampler = ClockSampler(local) if rank == 0 else None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host = time.perf_counter()
    for k in range(args.steps):
        step(k, last=(k == args.steps - 1))
    e1.record()
    host_enqueue_ms = (time.perf_counter() - t_host) * 1e3 / args.steps      # host time to enqueue a step (a floor for the step time)
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms = vr.sharding.max_over_ranks(e0.elapsed_time(e1), device="cuda")
    local_ms = vr.sharding.max_over_ranks(statistics.median(stage["total"]), device="cuda")
    if world > 1 and pipe.trace is not None:
        rows = pipe.trace_report()[-args.steps:]
        tot = [sum(r) for r in rows]
        slow = sorted(range(len(tot)), key=lambda i: -tot[i])[:4]
        sys.stderr.write("[c4 trace] rank %d: %d timed steps, sum %.3f ms (events %.3f ms), median step %.3f, slowest %s; last 6 (wait for inputs / stitch / signal): %s\n"
                         % (rank, len(rows), sum(tot), e0.elapsed_time(e1), statistics.median(tot), ["#%d %.3f" % (i, tot[i]) for i in slow],
                            " ".join("%.3f/%.3f/%.3f" % r for r in rows[-6:])))
    stats = [m.stats() for _, _, m in st.jobs]
    jobs_rank0 = [(j[0], list(j[1])) for j in st.jobs]
    exchange = None
    if args.verify and world > 1:                    # the two exchange steps timed on their own (diagnostic)
        def timed(fn, reps=10):
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return vr.sharding.max_over_ranks(a.elapsed_time(b) / reps, device="cuda")
        tok = torch.zeros(1, dtype=torch.int32, device="cuda")
        exchange = {"broadcast_ms": round(timed(lambda: vr.sharding.broadcast_frames(flats[0], 0)), 4),
                    ("frame_complete_allreduce_ms" if peer else "collect_ms"):
                        round(timed((lambda: vr.sharding.frame_complete(tok)) if peer else (lambda: vr.sharding.collect_shares(out, st.shares(), 0))), 4)}
    verified = None
    if args.verify and world > 1 and rank == 0:      # the assembled frame of the last step == both eyes stitched whole on this GPU
        one = vr.sharding.StereoRowBandStitcher(vr, tmpls, [in_size] * n, blend, gain, local, rank=0, world=1)
        ref = torch.zeros_like(out)
        one.stitch_local((full_ring if crop else ring)[(args.steps - 1) % RING], ref)
        torch.cuda.synchronize()
        verified = bool(torch.equal(ref, outs[(args.steps - 1) % 2]))
        del one, ref
    src_rows = [[int(m.src_rows()[c][1] - m.src_rows()[c][0]) for c in range(n)] for _, _, m in st.jobs]
    bcast_mb = bcast_bytes / 1e6
    crop_note = (": the source columns some mapper reads, %s of %d, cropped on rank 0 by one launch per step" % (sorted(set(w[1] for w in windows)), in_size[0])) if crop else ""
    st_split = st.split
    del st, pipe, outs, out, ring, flats, tmpls
    for pf in pfs:
        pf.close()
    torch.cuda.empty_cache()
    if not sub:
        pg_done()
    if rank != 0:
        return None
    peak, how = peaks()
    I = n * iw * ih * 3 // 2
    # both eyes are the same rig mirrored: P and roi area of eye 0 stand for both
    B = 2 * alg_bytes([in_size] * n, (W, He), stats[0]["pairs"], stats[0]["roi_area"], blend) - I
    ms_per_step = ms / args.steps
    blend_ms = statistics.median(stage["blend"])
    line = {
        "metric": "equirect output Mpix/s", "value": round(W * H * args.steps / (ms * 1e-3) / 1e6, 1), "unit": "Mpix/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 5), "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "frames_per_s": round(args.steps / (ms * 1e-3), 2),
        "config": {"workload": desc, "inputs": "%dx%dx%d I420 (octvr packed layout), ring of %d noise frames resident on rank 0" % (n, iw, ih, RING),
                   "output": "%dx%d 4:2:0 top-bottom (two %dx%d eyes)" % (W, H, W, He),
                   "l2": "working set per step exceeds the 126 MB L2; no flush",
                   "sharding": (("column" if st_split == "cols" else "row") + " bands of one stream: per step ONE NCCL broadcast of the %d input frames (%.1f MB%s) from rank 0 (issued one step "
                                "ahead, overlapping the previous stitch), (eye, band) stitch of the source rows the band reads "
                                "(%.1f MB frame in total)" % (n, bcast_mb, crop_note, W * H * 1.5 / 1e6)) if world > 1 else "single GPU, both eyes",
                   "assignment_rank0": jobs_rank0, "source_rows_converted_rank0": src_rows,
                   "band_output": ("stored by the blend kernels straight into rank 0's frame over NVLink peer memory (CUDA IPC); one 4-byte "
                                   "all-reduce per step = frame complete") if peer else ("NCCL send / recv to rank 0" if world > 1 else "local"),
                   "pairs_P_per_eye": stats[0]["pairs"], "roi_area_per_eye": stats[0]["roi_area"],
                   "template_build_s": round(t_tmpl, 2), "mapper_build_s": round(t_map, 2)},
        "alg_bytes_per_frame": int(B),
        "frac_of_hbm_roofline": {"whole_step_vs_measured_%.0f_x%d" % (peak, world): round(B / (ms_per_step * 1e-3) / 1e9 / (peak * world), 4)},
        "stage_ms_rank0": {k: round(statistics.median(v), 5) for k, v in stage.items()},
        "stitch_only_ms_max_over_ranks": round(local_ms, 5), "host_enqueue_ms_per_step_rank0": round(host_enqueue_ms, 4),
        "gpu_launches": sum(s_["launches_per_stitch"] for s_ in stats) * args.steps,
        "clocks": clocks,
    }
    if verified is not None:
        line["assembled_frame_equals_single_gpu_result"] = verified
    if exchange is not None:
        line["exchange_alone_ms"] = exchange
    if world == 1:
        ach = (B - I) / (blend_ms * 1e-3) / 1e9
        line["roofline"] = {"bound": "hbm", "kernel": "multiband stage, both eyes (k_mb_warp+k_mb_down+k_mb_band+k_mb_collapse+k_mb_final)",
                            "achieved": round(ach, 1), "peak": peak, "peak_source": how, "unit": "GB/s", "frac": round(ach / peak, 4),
                            "traffic": None, "alg_bytes_per_launch": int(B - I), "ms_per_launch": round(blend_ms, 5)}
    if sub:
        return line
    print(json.dumps(line))


def run_e2e(vr, util, tmpl, cfg, in_size, blend, gain, device, steps, world):
    """Same metric through AsyncMultiMapper.push/pop with HOST frames: H2D + stitch + D2H every step.  Headline figure: the
    caller's planes are page-locked (DMA straight from / to them); `pageable` repeats it with ordinary malloc'ed planes, as an
    ffmpeg filter would hand them over (the pipeline's threads stage them through pinned buffers, T1 / T5 of the reference)."""
    res = e2e_once(vr, util, tmpl, cfg, in_size, blend, gain, device, steps, world, pinned=True)
    try:
        pg = e2e_once(vr, util, tmpl, cfg, in_size, blend, gain, device, max(10, steps // 2), world, pinned=False)
        res["pageable"] = {k: pg[k] for k in ("value", "unit", "frames_per_s", "steps", "push_us_median")}
    except vr.OctvrError as ex:
        res["pageable"] = {"error": str(ex)}
    return res


def e2e_once(vr, util, tmpl, cfg, in_size, blend, gain, device, steps, world, pinned):
    import torch
    n = len(cfg["inputs"])
    iw, ih = in_size
    W, H = tmpl.out_size
    am = vr.AsyncMultiMapper([tmpl], [in_size] * n, (W, H), [blend], [0 if gain else -1], [(0.0, 0.0, 1.0, 1.0)], (0, 0), device=device)
    RING = 4

    def host(a):
        return torch.from_numpy(a).pin_memory().numpy() if pinned else np.ascontiguousarray(a)
    hin = []
    for k in range(RING):
        fr = []
        for c in range(n):
            buf = host(util.noise_frame(c, iw, ih, seed=99 + k))
            fr.append(util.i420_planes(buf, iw, ih))
        hin.append(fr)
    houts = [host(np.zeros(W * H * 3 // 2, np.uint8)) for _ in range(RING)]
    hout = [util.i420_planes(o, W, H) for o in houts]
    push_s = []
    depth = 3
    for k in range(depth):      # warm-up / fill
        am.push(hin[k % RING], hout[k % RING])
    for k in range(depth):
        am.pop()
    t0 = time.perf_counter()
    inflight = 0
    for k in range(steps):
        a0 = time.perf_counter()
        am.push(hin[k % RING], hout[k % RING])
        push_s.append(time.perf_counter() - a0)
        inflight += 1
        if inflight == depth:
            am.pop()
            inflight -= 1
    while inflight:
        am.pop()
        inflight -= 1
    dt = time.perf_counter() - t0
    am.close()
    dt = vr.sharding.max_over_ranks(dt, device="cuda")
    return {"value": round(W * H * steps * world / dt / 1e6, 1), "unit": "Mpix/s", "frames_per_s": round(steps * world / dt, 1),
            "h2d_bytes_per_step": n * iw * ih * 3 // 2, "d2h_bytes_per_step": W * H * 3 // 2,
            "api": "AsyncMultiMapper.push/pop (%s host planes, 3 frames in flight)" % ("pinned" if pinned else "pageable"), "steps": steps,
            "push_us_median": round(1e6 * statistics.median(push_s), 1)}


def cpu_baseline(workload, budget_s=20.0, steps=None, warmup=1):
    """The reference's CPU path on the same workload, on every host core.  kind "reference": oracle/_ref/libocvref.so, the
    UNMODIFIED reference CPU build behind oracle/refgen/ref_arm.cpp -- cv::cvtColor / cv::remap /
    GainCompensator / the octvr feather recipe or MultiBandBlender / cv::cvtColor, compiled from the reference's sources; kind
    "port": the C restatement oracle/liborc.so (OpenMP), used when the reference build did not travel (OCTVR_CPU_ARM=port
    forces it)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import refarm
    import util
    rig, blend, gain, desc = WORKLOADS[workload]
    if not isinstance(rig, str):
        raise SystemExit("the CPU arm runs the single-template workloads (c1, c2, c3); c4 is a GPU-only bench line")
    cfg, width, in_size = util.named_rig(rig)
    n = len(cfg["inputs"])
    iw, ih = in_size
    try:                                  # torchrun exports OMP_NUM_THREADS=1: the CPU arm is meant to use every host core
        O.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:                     # noqa: BLE001
        pass
    ot = O.build_template(cfg, width)
    use_ref = refarm.available() and os.environ.get("OCTVR_CPU_ARM", "reference") != "port"
    arm = None
    if use_ref:
        import tempfile
        try:
            with tempfile.TemporaryDirectory() as td:
                dat = os.path.join(td, "t.dat")
                O.dump_dat(ot, dat)       # "VRv11" file, byte-identical to the reference tool's (tests/test_oracle_golden.py)
                arm = refarm.RefArm(dat, in_size, blend, gain)
        except Exception as ex:           # noqa: BLE001  (a library that does not load on this box must not take the line with it)
            sys.stderr.write("reference CPU arm unavailable (%s): falling back to the C port\n" % ex)
            use_ref = False
    if use_ref:
        full = [util.noise_frame(c, iw, ih) for c in range(n)]
        run = lambda: arm.stitch(full)
        cores, kind = arm.threads(), "reference"
        note = ("reference CPU path = the unmodified reference CPU build (static libs compiled from the reference's own sources, SURVEY.md "
                "Appendix A) behind oracle/refgen/ref_arm.cpp: cvtColor -> remap -> GainCompensator -> feather / MultiBandBlender -> cvtColor, "
                "OpenCV's pthreads parallel_for_ + one thread per camera; built by oracle/build_ref.sh into oracle/_ref/")
    else:
        so = O.StitchOracle(ot, [in_size] * n, blend=blend, enable_gain=gain)
        frames = [util.i420_planes(util.noise_frame(c, iw, ih), iw, ih) for c in range(n)]
        run = lambda: so.stitch(frames)
        cores, kind = O.num_threads(), "port"
        note = ("reference CPU path = the C restatement under oracle/ (OpenMP over all host cores), pinned bit-exact against the unmodified "
                "reference build; oracle/_ref/libocvref.so (the reference build itself, oracle/build_ref.sh) was not available on this box")
    for _ in range(warmup):
        run()
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
        if steps is not None and (len(times) >= steps or time.perf_counter() - t_start > budget_s):
            break
        if steps is None and (time.perf_counter() - t_start > budget_s or len(times) >= 10):
            break
    W, H = ot.out_size
    med = statistics.median(times)
    return {"value": round(W * H / med / 1e6, 2), "unit": "Mpix/s", "cores": cores, "kind": kind,
            "frames_per_s": round(1.0 / med, 3), "ms_per_frame": round(med * 1e3, 1), "steps_timed": len(times),
            "sample": "%d full frames of %s (median), after %d warm-up" % (len(times), desc, warmup),
            "note": note}


def cpu_baseline_fast(width=None):
    """FastMapper on the host: the numpy restatement of mapper_fast.cpp + remap_weighted.cl (oracle.FastMapperOracle, one thread)
    on ONE frame of the `fast` workload.  The reference's own FastMapper needs an OpenCL device, so there is nothing else to time."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import util
    rig, _, _, desc = WORKLOADS["fast"]
    cfg, w0, in_size = util.named_rig(rig)
    width = width or w0
    n = len(cfg["inputs"])
    ot = O.build_template(cfg, width, use_roi=False, with_seams=False)
    fo = O.FastMapperOracle(ot, [in_size] * n)
    frames = [O.fast_noise_frame(c, *in_size) for c in range(n)]
    t0 = time.perf_counter()
    with np.errstate(over="ignore"):
        fo.stitch_nv12(frames)
    dt = time.perf_counter() - t0
    W, H = ot.out_size
    return {"value": round(W * H / dt / 1e6, 2), "unit": "Mpix/s", "cores": 1, "kind": "port", "frames_per_s": round(1.0 / dt, 3),
            "ms_per_frame": round(dt * 1e3, 1), "steps_timed": 1, "sample": "1 full frame of %s at output width %d" % (desc, W),
            "note": "numpy restatement of vr::FastMapper (oracle/oracle.py); the reference's FastMapper itself needs an OpenCL device"}


def run_reference(args):
    """The reference's CPU path on the box's host cores, on the headline arm's workload / metric.  A step = one full output frame
    (the CPU port needs ~0.1 s for a C2 frame on 16 threads, so --steps K is honoured as given up to a wall-clock budget of
    ~3 minutes; the line says how many steps were actually timed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rig, blend, gain, desc = WORKLOADS[args.workload]
    warm = max(1, min(args.warmup, 3))
    cb = cpu_baseline(args.workload, steps=max(1, args.steps), warmup=warm, budget_s=180.0)
    line = {"impl": "reference", "metric": "equirect output Mpix/s", "value": cb["value"], "unit": "Mpix/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": cb["steps_timed"], "warmup": warm,
            "ms_per_step": cb["ms_per_frame"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": {"workload": desc, "note": cb["note"]},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sub", action="store_true", help="c2: skip the c3 / c5 / c4_rowband sub-benchmarks")
    ap.add_argument("--concurrency", type=int, default=1, help="c5: CUDA streams (and Mappers) a rank serves its video streams on "
                    "(measured on B200: 2-4 are 5 %% slower per frame than 1)")
    ap.add_argument("--verify", action="store_true", help="c4, N > 1: rank 0 also stitches the last frame whole and compares")
    ap.add_argument("--rowband", action="store_true", help="N > 1: all ranks stitch ONE stream, split by output row bands "
                    "(NCCL broadcast of the inputs + band collection inside the timed region; strong scaling)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c4":
        run_stereo(args)
    elif args.workload == "c5":
        run_streams(args)
    elif args.workload == "fast":
        line = run_fast(args)
        if line is not None:
            if not args.no_cpu and int(os.environ.get("WORLD_SIZE", "1")) == 1:
                line["cpu_baseline"] = cpu_baseline_fast()
            print(json.dumps(line))
    elif args.rowband and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        run_rowband(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
