#!/usr/bin/env python
"""bench.py -- headline benchmark of the go-webp pixel hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 1|2|3|4]

Headline workload (BASELINE.json configs[1], --config 1): a batch of 256 synthetic
1920x1080 RGBA frames per GPU, RGBA -> YUV420(+A) import.  A "step" is one pass of the
import kernel over the resident batch (one kernel launch).  Weak scaling: every rank owns
its own 256-frame batch, no data-path collective (images are independent).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs in HBM
before the timed region); `e2e` is the same metric through the drop-in batched entry point
picture.ImportRGBABatch (-> b2_import_batch_staged) with pinned HOST buffers, H2D and D2H
inside the timed region, next to the box's own copy ceiling measured in the same run.
`roofline` is for the import kernel; `cpu_baseline` is the CPU oracle port timed on this
box's host cores.  `other_configs` carries, at EVERY N and from every rank's share, the
other BASELINE configs: [2] VP8L chains on 64 x 4096^2, [3] 1024 x 4K fancy upsample sharded
by image (strong scaling), [4] the mixed-size batch, LPT by calibrated time -- each with its
own device-resident and end-to-end figure.  --config k makes config k the headline.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 1920, 1080
SEED = 0xB2000001  # config k uses seed 0xB2000000 + k
METRIC = "Mpixel/s batched 1080p RGBA<->YUV420 + VP8L inverse at 1-8 B200; % of HBM peak"
ALPHA_MODES = {"opaque": 1, "sparse": 2, "binary": 3, "raw": 0}
# algorithmic bytes per pixel: 4 in + 1 Y + 0.5 UV + 1 A
IMPORT_BYTES_PER_PX = 4 + 1 + 0.5 + 1


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons while the GPU is under load."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, device_index: int):
        super().__init__(daemon=True)
        self.samples = []  # (t, sm_mhz, reasons_bitmask)
        self.device_index = device_index
        self.stop_flag = False
        self.ok = False
        self.sm_max = None
        try:
            if os.environ.get("B2_CLOCKS") == "smi":   # exercise the nvidia-smi fallback
                raise RuntimeError("B2_CLOCKS=smi")
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = device_index
            if vis:
                try:
                    idx = int(vis.split(",")[device_index])
                except Exception:
                    idx = device_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:   # fall back to polling nvidia-smi (the profiling recipe's clocks line)
            self.ok = False
            self.err = repr(e)
            self.smi_index = device_index

    def run_smi(self):
        import subprocess
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.smi_index
        if vis:
            try:
                idx = int(vis.split(",")[self.smi_index])
            except Exception:
                pass
        try:
            proc = subprocess.Popen(["nvidia-smi", "-i", str(idx), "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                                     "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        try:
            for line in proc.stdout:
                if self.stop_flag:
                    break
                f = [x.strip() for x in line.split(",")]
                try:
                    self.samples.append((time.perf_counter(), int(f[0]), int(f[2], 16)))
                    self.sm_max = int(f[1])
                    self.ok = True
                except Exception:
                    pass
        finally:
            proc.kill()

    def count_between(self, t0, t1):
        return sum(1 for s in self.samples if t0 <= s[0] <= t1)

    def run(self):
        if not self.ok:
            self.run_smi()
            return
        nv = self.nv
        fails = 0
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, int(r)))
            except Exception:
                fails += 1
                if fails > 20 and not self.samples:   # NVML answers nothing here: poll nvidia-smi instead
                    self.smi_index = self.device_index
                    self.ok = False
                    self.run_smi()
                    return
            time.sleep(0.001)

    def summary(self, t0, t1, t_load0, extra=None):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0, "window": "unavailable"}
        win = [s for s in self.samples if t0 <= s[0] <= t1]
        window = "timed"
        if len(win) < 3:
            win = [s for s in self.samples if t_load0 <= s[0] <= t1]
            window = "warmup+timed"
        if len(win) < 3 and extra is not None:
            win = [s for s in self.samples if extra[0] <= s[0] <= extra[1]]
            window = "same kernel re-run untimed right after the timed region (too few samples inside it)"
        if not win:
            win, window = self.samples, "run"
        bits = 0
        for s in win:
            bits |= s[2]
        reasons = sorted(n for b, n in self.REASONS.items() if bits & b)
        return {"sm_mhz": float(np.median([s[1] for s in win])), "sm_max_mhz": self.sm_max,
                "reasons": reasons, "samples": len(win), "window": window}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---- CPU side (oracle port): used only as the reported baseline -------------------
def cpu_import_rate(n_frames: int, threads: int, alpha_mode: int, scalar: bool = False, min_seconds: float = 0.0):
    """Times the oracle's RGBA->YUV420A import over n_frames 1080p frames with
    'threads' host threads (one image per task); with min_seconds, one untimed pass and then
    whole passes until that much wall time has gone by.  Returns (Mpx/s, seconds)."""
    import oracle
    L = oracle.lib(scalar)
    distinct = min(n_frames, 32)
    src = np.empty((distinct, H, W), np.uint32)
    for i in range(distinct):
        src[i] = oracle.synth_u32(SEED, i, 0, W, H, alpha_mode)
    if distinct < n_frames:
        src = np.concatenate([src] * ((n_frames + distinct - 1) // distinct))[:n_frames]
    src = np.ascontiguousarray(src)
    uvw, uvh = (W + 1) // 2, (H + 1) // 2
    y = np.empty((n_frames, H, W), np.uint8)
    u = np.empty((n_frames, uvh, uvw), np.uint8)
    v = np.empty((n_frames, uvh, uvw), np.uint8)
    a = np.empty((n_frames, H, W), np.uint8)
    ha = np.zeros(n_frames, np.int32)
    vp = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    run = lambda: L.orc_import_batch_mt(vp(src), n_frames, W, H, vp(y), vp(u), vp(v), vp(a), vp(ha), 0, threads)  # noqa: E731
    if min_seconds <= 0.0:
        t0 = time.perf_counter()
        rc = run()
        dt = time.perf_counter() - t0
        assert rc == 0
        return n_frames * W * H / dt / 1e6, dt
    assert run() == 0          # untimed: first touch of the output pages
    passes, t0 = 0, time.perf_counter()
    while True:
        assert run() == 0
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return passes * n_frames * W * H / dt / 1e6, dt


def libwebp_import_rate(n_frames: int, threads: int, alpha_mode: int):
    """Upstream libwebp 1.6.0 (SIMD C) WebPPictureImportRGBA, if Pillow's bundled
    binary is present: the 'upstream SIMD C' figure of BASELINE.md 3."""
    try:
        import oracle
        from concurrent.futures import ThreadPoolExecutor
        from oracle import libwebp_ref as ref
        ref.lib()
        frames = [oracle.synth_u32(SEED, i, 0, W, H, alpha_mode).view(np.uint8) for i in range(min(8, n_frames))]

        def work(i):
            ref.import_yuva(frames[i % len(frames)], W, H, "RGBA")
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, range(n_frames)))
        dt = time.perf_counter() - t0
        return n_frames * W * H / dt / 1e6
    except Exception:
        return None


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


class Dist:
    """torch.distributed plumbing: barrier + max / sum / gather over ranks (NCCL; no data-path use)."""

    def __init__(self):
        self.rank, self.world, self.local = dist_env()
        import torch
        self.torch = torch
        torch.cuda.set_device(self.local)
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _red(self, x, op):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._red(x, self.dist.ReduceOp.MAX if self.world > 1 else None)

    def sum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM if self.world > 1 else None)

    def gather(self, x):
        if self.world == 1:
            return [float(x)]
        t = self.torch.zeros(self.world, dtype=self.torch.float64, device="cuda")
        t[self.rank] = float(x)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def bcast_list(self, values):
        """rank 0's list of floats to everyone (the calibrated cost model of configs[4])."""
        if self.world == 1:
            return list(values)
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        self.dist.broadcast(t, 0)
        return [float(v) for v in t.tolist()]

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed_device(ctx, D, run, steps, warm=3, slots=(6, 7)):
    """ms per call of run() on stream 0: CUDA events, barrier + synchronize on both sides, max over ranks."""
    for _ in range(warm):
        run()
    D.barrier()
    ctx.event_record(slots[0])
    for _ in range(steps):
        run()
    ctx.event_record(slots[1])
    D.barrier()
    return D.max(ctx.elapsed_ms(slots[0], slots[1]) / steps)


def timed_wall(D, run, steps, warm=1):
    """seconds per call of a synchronous host-level call (the end-to-end entry points), max over ranks."""
    for _ in range(warm):
        run()
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    D.barrier()
    return D.max((time.perf_counter() - t0) / steps)


def workload_config(args):
    return {"workload": f"configs[1]: {args.frames} x {W}x{H} synthetic RGBA frames per GPU, RGBA->YUV420A import "
                        f"(Y,U,V,A planes + has_alpha flag), alpha mix '{args.alpha}'",
            "frames_per_gpu": args.frames, "width": W, "height": H, "alpha_mix": args.alpha,
            "partition": "images sharded by rank, no collective",
            "l2": "input batch (2.1 GB) and outputs (1.3 GB) are >> 126 MB L2; no flush needed",
            "seed": hex(SEED)}


# ---- BASELINE configs[2], [3], [4]: every rank's share, device-resident and end to end ------------
def config3_upsample(ctx, D, peak, steps, total_frames=1024, e2e_frames=32):
    """configs[3]: 1024 x 3840x2160 YUV420 -> RGBA fancy upsampling sharded by image over the ranks
    (strong scaling: 1024 / N frames per GPU).  e2e = picture.YUVAToRGBABatch on pinned planes."""
    from go_webp_b200 import _abi, picture
    from go_webp_b200.partition import contiguous_shard
    from go_webp_b200.pipeline import UpsamplePipeline
    w, h = 3840, 2160
    mine = contiguous_shard(total_frames, D.world, D.rank)
    n = len(mine)
    up = UpsamplePipeline(ctx, n, w, h, _abi.MODE_RGBA, host=False)
    up.fill_synthetic(0xB2000004, first_img=mine.start)
    ms = timed_device(ctx, D, up.run_device, max(3, min(steps, 10)))
    up.close()
    px = total_frames * w * h
    out = {"workload": f"configs[3]: {total_frames} x {w}x{h} YUV420 -> RGBA fancy upsample, {n} frames on this rank "
                       f"(image-sharded, strong scaling)", "n_gpus": D.world, "scaling": "strong",
           "value": px / ms / 1e3, "unit": "Mpixel/s", "ms_per_step": ms,
           "roofline": {"bound": "hbm", "bytes_per_px": 5.5, "achieved": n * w * h * 5.5 / ms / 1e6, "peak": peak,
                        "frac": n * w * h * 5.5 / ms / 1e6 / peak, "unit": "GB/s", "per": "GPU"}}
    # end to end through the drop-in entry point: pinned Y/U/V in, pinned RGBA out, a bounded sample per rank
    m = min(e2e_frames, n)
    pics = [picture.Picture(Width=w, Height=h) for _ in range(m)]
    picture.AllocResidentBatch(pics, "yuva", ctx)
    for i, pic in enumerate(pics):   # synthetic planes straight into the pinned mirrors, then device copies dropped
        mem = pic.memory_
        ctx.synth_fill(mem.dev("y"), mem.ys // 4, h, mem.ys, mem.ys * h, 1, 0xB2000004, mine.start + i, 1)
        ctx.synth_fill(mem.dev("u"), mem.uvs // 4, mem.uvh, mem.uvs, mem.uvs * mem.uvh, 1, 0xB2000004, mine.start + i, 2)
        ctx.synth_fill(mem.dev("v"), mem.uvs // 4, mem.uvh, mem.uvs, mem.uvs * mem.uvh, 1, 0xB2000004, mine.start + i, 3)
        mem.dev_valid = True
        mem.download(with_alpha=False)
    dst = picture.PinnedBytes(m * 4 * w * h, ctx)
    dsts = [dst.array[i * 4 * w * h:(i + 1) * 4 * w * h].reshape(h, 4 * w) for i in range(m)]

    def run():
        for pic in pics:
            pic.memory_.dev_valid = False     # host planes are the source of truth: upload every step
        assert picture.YUVAToRGBABatch(pics, _abi.MODE_RGBA, dsts, ctx, chunk=4) is not None
    for pic in pics:
        pic.ColorSpace = picture.WEBP_YUV420
    t = timed_wall(D, run, 2)
    agg_px = D.sum(m * w * h)
    out["e2e"] = {"value": agg_px / t / 1e6, "unit": "Mpixel/s", "api": "picture.YUVAToRGBABatch (b2_upsample_batch_staged)",
                  "sample": f"{m} frames per rank per step", "h2d_bytes_per_step": int(m * w * h * 1.5),
                  "d2h_bytes_per_step": m * w * h * 4}
    for pic in pics:
        picture.WebPPictureFree(pic)
    return out


def config2_vp8l(ctx, D, peak, steps, total_images=64, e2e_images=16):
    """configs[2]: VP8L inverse chains on 64 x 4096x4096 ARGB images sharded over the ranks: fused
    cross-colour^-1 + add-green (32x32 tiles), 8-bit and 2-bit colour indexing, and the full
    predictor + cross-colour + add-green chain.  e2e = b2_vp8l_batch_staged with pinned buffers."""
    from go_webp_b200.partition import contiguous_shard
    from go_webp_b200.pipeline import ArgbPipeline
    w = h = 4096
    mine = contiguous_shard(total_images, D.world, D.rank)
    n = len(mine)
    out = {"workload": f"configs[2]: {total_images} x {w}x{h} ARGB, {n} images on this rank", "n_gpus": D.world,
           "scaling": "strong", "kernels": {}}
    k = max(3, min(steps, 10))
    for op, bits, bpp, name in [("inverse_chain", 5, 8.0039, "crosscolor_inv_addgreen_fused"),
                                ("color_index_inverse", 0, 8.0, "color_index_8bit"),
                                ("color_index_inverse", 2, 5.0, "color_index_2bit"),
                                ("predictor_inverse", 5, 8.0039, "predictor_inverse")]:
        ap = ArgbPipeline(ctx, n, w, h, op, bits, host=(name == "crosscolor_inv_addgreen_fused"), host_images=e2e_images)
        ap.fill_synthetic(0xB2000003, first_img=mine.start)
        ms = timed_device(ctx, D, ap.run_device, k if op != "predictor_inverse" else 3)
        px = total_images * w * h
        out["kernels"][name] = {"mpixel_per_s": px / ms / 1e3, "ms_per_step": ms, "bytes_per_px": bpp,
                                "frac": n * w * h * bpp / ms / 1e6 / peak}
        if ap.host:
            m = ap.host_images
            t = timed_wall(D, ap.run, 2)
            out["e2e"] = {"value": D.sum(m * w * h) / t / 1e6, "unit": "Mpixel/s",
                          "api": "b2_vp8l_batch_staged(B2_VP8L_INVERSE_CHAIN) = dsp.VP8LInverseTransformBatch's call, pinned buffers",
                          "sample": f"{m} images per rank per step", "h2d_bytes_per_step": ap.h2d_bytes, "d2h_bytes_per_step": ap.d2h_bytes}
        ap.close()
    # the whole chain libwebp's encoder writes (cross-colour, predictor, subtract-green), one pass vs three launches
    from go_webp_b200.pipeline import ChainPipeline
    cp = ChainPipeline(ctx, n, w, h, 5)
    cp.fill_synthetic(0xB2000003, first_img=mine.start)
    ms_f = timed_device(ctx, D, cp.run_device, 3)
    ms_u = timed_device(ctx, D, cp.run_unfused, 3)
    px = total_images * w * h
    out["kernels"]["full_chain_fused_one_pass"] = {"mpixel_per_s": px / ms_f / 1e3, "ms_per_step": ms_f, "bytes_per_px": 8.008,
                                                   "frac": n * w * h * 8.008 / ms_f / 1e6 / peak}
    out["kernels"]["full_chain_three_launches"] = {"mpixel_per_s": px / ms_u / 1e3, "ms_per_step": ms_u, "bytes_per_px": 24.008,
                                                   "speedup_of_fused": ms_u / ms_f}
    cp.close()
    kk = out["kernels"]["crosscolor_inv_addgreen_fused"]
    out.update({"value": kk["mpixel_per_s"], "unit": "Mpixel/s", "ms_per_step": kk["ms_per_step"],
                "roofline": {"bound": "hbm", "bytes_per_px": 8.0039, "frac": kk["frac"], "peak": peak, "unit": "GB/s", "per": "GPU"}})
    return out


def config4_mixed(ctx, D, peak, steps, n_images=256):
    """configs[4]: 256 images, 64x64 .. 8192x8192 with odd widths / heights, full pixel pipeline (import ->
    fancy upsample -> VP8L inverse chain), LPT-partitioned over the ranks by CALIBRATED TIME: one warm-up
    pass times each (width class, size class) group of the batch, cost = bytes / rate(group)."""
    from go_webp_b200 import partition
    from go_webp_b200.pipeline import MixedPipeline, calibrate_mixed_rates, mixed_sizes
    sizes = mixed_sizes(n_images, 0xB2000005)
    groups = [partition.size_class(w_, h_) for (w_, h_) in sizes]
    rates = calibrate_mixed_rates(ctx, sizes) if D.rank == 0 else {}
    keys = sorted(set(groups))
    vals = D.bcast_list([rates.get(g, 0.0) for g in keys])
    rates = dict(zip(keys, vals))
    costs = partition.calibrated_costs(sizes, rates)
    k = max(3, min(steps, 10))
    rounds, best = (5 if D.world > 1 else 1), None
    for it in range(rounds):   # the class model first, then the measured per-GPU times of a pass correct it (every rank computes the same update)
        parts = partition.lpt_partition(costs, D.world)
        p = MixedPipeline(ctx, [sizes[i] for i in parts[D.rank]])
        p.fill_synthetic(0xB2000005)
        for _ in range(3):
            p.run_device()
        D.barrier()
        ctx.event_record(8)
        for _ in range(k):
            p.run_device()
        ctx.event_record(9)
        D.barrier()
        per_gpu = D.gather(ctx.elapsed_ms(8, 9) / k)
        res = {"ms": max(per_gpu), "per_gpu": per_gpu, "px": D.sum(p.pixels), "by": D.sum(p.algorithmic_bytes()), "round": it}
        p.close()
        if best is None or res["ms"] < best["ms"]:
            best = res
        pred = [sum(costs[i] for i in part) for part in parts]
        norm = sum(per_gpu) / sum(pred)
        for part, pr, me in zip(parts, pred, per_gpu):
            f = (me / (pr * norm)) ** 0.7
            for i in part:
                costs[i] *= f
    ms, per_gpu, px, by = best["ms"], best["per_gpu"], best["px"], best["by"]
    out = {"workload": f"configs[4]: {n_images} mixed-size images (64x64..8192x8192, odd sizes), import -> fancy upsample -> "
                       f"cross-colour^-1 + add-green, 3 launches per pass over the rank's share", "n_gpus": D.world,
           "scaling": "strong", "value": px / ms / 1e3, "unit": "Mpixel/s", "ms_per_step": ms,
           "per_gpu_ms": per_gpu, "imbalance_max_over_mean": ms * D.world / sum(per_gpu),
           "partition": "LPT on calibrated time (bytes / measured rate of the image's width-alignment x size class), then four "
                        "rounds in which the measured per-GPU times of a pass correct the costs; the best of the five "
                        f"partitions (round {best['round']}) is the one reported, all timed alike",
           "roofline": {"bound": "hbm", "achieved": by / D.world / ms / 1e6, "peak": peak, "unit": "GB/s",
                        "frac": by / D.world / ms / 1e6 / peak, "per": "GPU (mean share / slowest rank)"}}
    return out


def config1_import(ctx, D, args, peak, peak_src, sampler):
    from go_webp_b200 import picture
    from go_webp_b200.pipeline import ImportPipeline
    alpha_mode = ALPHA_MODES[args.alpha]
    rank, world = D.rank, D.world
    pipe = ImportPipeline(ctx, args.frames, W, H, "RGBA", import_alpha=True, chunk=16, host=True)
    pipe.fill_synthetic(SEED, first_img=rank * args.frames, alpha_mode=alpha_mode, to_host=True)
    t_load0 = time.perf_counter()
    for _ in range(args.warmup):
        pipe.run_device()
    D.barrier()
    launches0 = ctx.launches
    t0 = time.perf_counter()
    ctx.event_record(0)
    for _ in range(args.steps):
        pipe.run_device()
    ctx.event_record(1)
    D.barrier()
    t1 = time.perf_counter()
    launches = ctx.launches - launches0
    ms = D.max(ctx.elapsed_ms(0, 1))
    extra = None
    if sampler.count_between(t_load0, t1) < 3:   # keep the same kernel running (untimed) until the sampler has seen it
        e0 = time.perf_counter()
        while time.perf_counter() - e0 < 0.5:
            for _ in range(10):
                pipe.run_device()
            ctx.sync()
        extra = (e0, time.perf_counter())
    clocks = sampler.summary(t0, t1, t_load0, extra)
    px_step = world * args.frames * W * H
    ms_step = ms / args.steps
    value = px_step / ms_step / 1e3  # Mpixel/s
    gbs_rank = args.frames * W * H * IMPORT_BYTES_PER_PX / ms_step / 1e6

    # sustained: the same launch back to back for >= 2 s, with the clock record of that window
    sustained = None
    if not args.no_extras:
        n_loop = int(2200.0 / ms_step) + 1
        D.barrier()
        ts0 = time.perf_counter()
        ctx.event_record(2)
        for _ in range(n_loop):
            pipe.run_device()
        ctx.event_record(3)
        D.barrier()
        ts1 = time.perf_counter()
        sms = D.max(ctx.elapsed_ms(2, 3)) / n_loop
        sustained = {"seconds": sms * n_loop / 1e3, "launches": n_loop, "ms_per_step": sms, "value": px_step / sms / 1e3,
                     "frac": args.frames * W * H * IMPORT_BYTES_PER_PX / sms / 1e6 / peak,
                     "clocks": sampler.summary(ts0, ts1, ts0)}

    # parity spot check against the oracle (outside every timed region)
    parity = "skipped"
    if rank == 0:
        import oracle
        oracle.build()
        ok = True
        for i in (0, args.frames - 1):
            src = oracle.synth_u32(SEED, rank * args.frames + i, 0, W, H, alpha_mode).view(np.uint8)
            ey, eu, ev, ea, eha = oracle.import_yuva(src, W, H, "RGBA")
            gy, gu, gv, ga, gha = pipe.device_planes(i)
            ok &= bool((gy == ey).all() and (gu == eu).all() and (gv == ev).all() and (ga == ea).all() and gha == eha)
        parity = "bit-exact vs oracle on frames 0 and %d" % (args.frames - 1) if ok else "MISMATCH"

    # end to end through the drop-in entry point: pinned host sources -> picture.ImportRGBABatch -> pinned planes
    e2e_steps = max(2, min(args.steps, 8))
    pics = [picture.Picture(Width=W, Height=H) for _ in range(args.frames)]
    srcs = [pipe.src_view(i) for i in range(args.frames)]
    strides = [4 * W] * args.frames
    s0 = ctx.copy_stats()
    run_api = lambda: picture.ImportRGBABatch(pics, srcs, strides, "RGBA", ctx, chunk=16)  # noqa: E731
    assert run_api() == 1
    s1 = ctx.copy_stats()
    te = timed_wall(D, run_api, e2e_steps, warm=1)
    e2e_value = px_step / te / 1e6
    tp = timed_wall(D, pipe.run, e2e_steps, warm=1)           # the bench-shaped twin, for comparison
    tc = timed_wall(D, pipe.run_copies_only, e2e_steps, warm=1)  # the same copies with no kernel: this box's ceiling
    h2d, d2h = s1["h2d_bytes"] - s0["h2d_bytes"], s1["d2h_bytes"] - s0["d2h_bytes"]
    if rank == 0 and parity != "MISMATCH":
        y, u, v, a, ha = pipe.planes(0)
        gy, gu, gv, ga, gha = pipe.device_planes(0)
        p0 = pics[0]
        if not ((y == gy).all() and (u == gu).all() and (v == gv).all() and ha == gha and (p0.Y == gy).all()
                and (p0.U == gu).all() and (p0.V == gv).all() and (p0.A is not None) == gha and (not gha or (p0.A == ga).all())):
            parity = "MISMATCH (e2e)"
    ceiling_gbs = world * (pipe.h2d_bytes + pipe.d2h_bytes) / tc / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "go_webp_b200.picture.ImportRGBABatch (drop-in batched entry point -> "
                                           "b2_import_batch_staged; pinned sources, device-resident pictures with pinned mirrors)",
                "pipeline_value": px_step / tp / 1e6, "pipeline_api": "pipeline.ImportPipeline.run (bench-shaped twin)",
                "gbs": world * (h2d + d2h) / te / 1e9,
                "ceiling_gbs": ceiling_gbs, "frac_of_ceiling": world * (h2d + d2h) / te / 1e9 / ceiling_gbs,
                "ceiling": "the same H2D + D2H copies of this step with no kernel, all ranks at once (live; "
                           "profiles/r02_pcie_ceiling.md has the bare-CUDA microbenchmark of the box)"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": gbs_rank, "peak": peak, "unit": "GB/s",
                     "frac": gbs_rank / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                     "kernel": "import_fast_kernel<RGBA,alpha>", "algorithmic_bytes_per_px": IMPORT_BYTES_PER_PX,
                     "frac_of_8000_nominal": gbs_rank / 8000.0, "per": "GPU (rank 0 step time = max over ranks)"},
        "parity": parity,
    }
    if sustained:
        line["sustained"] = sustained
    for pic in pics:
        picture.WebPPictureFree(pic)
    return line, pipe


def other_kernels(ctx, peak, steps):
    """Device-resident numbers for the rest of the path at reduced batch sizes
    (explanatory; not the headline): 4K fancy upsample and the VP8L transforms."""
    from go_webp_b200 import _abi
    from go_webp_b200.pipeline import ArgbPipeline, ImportPipeline, UpsamplePipeline
    out = {}

    def timed(run, n_steps):
        for _ in range(3):
            run()
        ctx.sync()
        ctx.event_record(2)
        for _ in range(n_steps):
            run()
        ctx.event_record(3)
        ctx.sync()
        return ctx.elapsed_ms(2, 3) / n_steps

    k = max(3, min(steps, 20))
    for name, frames, w, h, mode, wa, bpp in [
            ("fancy_upsample_4k_rgba", 64, 3840, 2160, _abi.MODE_RGBA, False, 5.5),
            ("fancy_upsample_4k_bgra_with_alpha_plane", 32, 3840, 2160, _abi.MODE_BGRA, True, 6.5),
            ("fancy_upsample_4k_rgb", 32, 3840, 2160, _abi.MODE_RGB, False, 4.5),
            ("fancy_upsample_1080p_rgba", 256, 1920, 1080, _abi.MODE_RGBA, False, 5.5)]:
        up = UpsamplePipeline(ctx, frames, w, h, mode, with_alpha=wa, host=False)
        up.fill_synthetic(0xB2000004)
        ms = timed(up.run_device, k)
        px = frames * w * h
        out[name] = {"frames": frames, "mpixel_per_s": px / ms / 1e3, "bytes_per_px": bpp,
                     "gbs": px * bpp / ms / 1e6, "frac": px * bpp / ms / 1e6 / peak}
        up.close()
    # 3-byte imports (WebPPictureImportRGB / BGR: 3 B in, 1.5 B out per pixel), 128 x 1080p
    al = lambda v: (v + 255) // 256 * 256  # noqa: E731
    n3 = 128
    s3, y3 = ctx.alloc(n3 * al(3 * W * H)), ctx.alloc(n3 * al(W * H))
    u3, v3 = ctx.alloc(n3 * al(W * H // 4)), ctx.alloc(n3 * al(W * H // 4))
    ctx.synth_fill(s3.ptr, (3 * W) // 4, H, 3 * W, al(3 * W * H), n3, SEED, 0, 4)
    ctx.sync()
    j3 = (_abi.ImportJob * n3)(*[_abi.ImportJob(s3.ptr + i * al(3 * W * H), y3.ptr + i * al(W * H), u3.ptr + i * al(W * H // 4),
                                                 v3.ptr + i * al(W * H // 4), None, 3 * W, W, W // 2, 0, W, H) for i in range(n3)])
    ms = timed(lambda: ctx._chk(ctx.lib.b2_rgba_to_yuv420a(ctx.h, 0, j3, n3, _abi.ORDER_RGB, 0, None)), k)
    out["import_rgb_3_bytes_1080p"] = {"frames": n3, "mpixel_per_s": n3 * W * H / ms / 1e3, "bytes_per_px": 4.5,
                                       "gbs": n3 * W * H * 4.5 / ms / 1e6, "frac": n3 * W * H * 4.5 / ms / 1e6 / peak}
    for b in (s3, y3, u3, v3):
        b.free()
    # picture pre-processing on the planes an import just wrote: binary alpha makes every
    # 8x8 block partially transparent (SmoothenBlock rewrites its luma: read A + Y, write Y = 3 B/px);
    # blend reads and writes Y, U, V, A (5 B/px on the first pass)
    ip = ImportPipeline(ctx, 128, W, H, host=False)
    ip.fill_synthetic(SEED, 0, ALPHA_MODES["binary"])
    ip.run_device()
    ctx.sync()
    pj = (_abi.PictureJob * 128)(*[_abi.PictureJob(j.y, j.u, j.v, j.a, None, j.y_stride, j.uv_stride, j.a_stride, 0, j.width, j.height)
                                   for j in ip.jobs])
    for name, bpp, fn in [("picture_cleanup_transparent_1080p_yuva", 3.0,
                           lambda: ctx._chk(ctx.lib.b2_picture_cleanup_transparent(ctx.h, 0, pj, 128, 0))),
                          ("picture_blend_alpha_1080p_yuva", 5.0,
                           lambda: ctx._chk(ctx.lib.b2_picture_blend_alpha(ctx.h, 0, pj, 128, 0, 0x336699)))]:
        ms = timed(fn, k)
        px = 128 * W * H
        out[name] = {"frames": 128, "mpixel_per_s": px / ms / 1e3, "bytes_per_px": bpp,
                     "gbs": px * bpp / ms / 1e6, "frac": px * bpp / ms / 1e6 / peak}
    ip.close()
    for op, bits, bpp, name in [("inverse_chain", 5, 8.0039, "vp8l_crosscolor_inv_addgreen_fused"),
                                ("forward_chain", 5, 8.0039, "vp8l_subgreen_crosscolor_fwd_fused"),
                                ("add_green", 0, 8.0, "vp8l_add_green"),
                                ("color_index_inverse", 0, 8.0, "vp8l_color_index_8bit"),
                                ("color_index_inverse", 1, 6.0, "vp8l_color_index_4bit"),
                                ("color_index_inverse", 2, 5.0, "vp8l_color_index_2bit"),
                                ("color_index_inverse", 3, 4.5, "vp8l_color_index_1bit"),
                                ("predictor_forward", 5, 8.0039, "vp8l_predictor_forward"),
                                ("predictor_inverse", 5, 8.0039, "vp8l_predictor_inverse_wavefront")]:
        ap = ArgbPipeline(ctx, 64, 4096, 4096, op, bits)   # configs[2]: 64 x 4096^2
        ap.fill_synthetic(0xB2000003)
        ms = timed(ap.run_device, k)
        px = 64 * 4096 * 4096
        out[name] = {"frames": 64, "mpixel_per_s": px / ms / 1e3, "bytes_per_px": bpp,
                     "gbs": px * bpp / ms / 1e6, "frac": px * bpp / ms / 1e6 / peak}
        ap.close()
    from go_webp_b200.pipeline import PalettePipeline
    for psize, bpp, name in [(256, 8.0, "vp8l_apply_palette_256_colours_8bit"), (16, 4.5, "vp8l_apply_palette_16_colours_4bit"),
                             (4, 4.25, "vp8l_apply_palette_4_colours_2bit"), (2, 4.125, "vp8l_apply_palette_2_colours_1bit")]:
        pp = PalettePipeline(ctx, 64, 4096, 4096, psize)   # colour -> packed index: 4 B/px in, 4 / 2^xbits out
        pp.fill_synthetic(0xB2000003)
        ms = timed(pp.run_device, k)
        px = 64 * 4096 * 4096
        out[name] = {"frames": 64, "mpixel_per_s": px / ms / 1e3, "bytes_per_px": bpp,
                     "gbs": px * bpp / ms / 1e6, "frac": px * bpp / ms / 1e6 / peak}
        pp.close()
    # the 8f-4 rows on 128 alpha / luma planes of 1080p (2 B/px; level smoothing 3: its level scan reads the plane too)
    n, st = 128, W
    src, dst = ctx.alloc(n * st * H), ctx.alloc(4 * n * st * H)
    ctx.synth_fill(src.ptr, st // 4, H, st, st * H, n, SEED + 7, 0, 4)
    ctx.sync()
    k5 = max(3, min(steps, 5))

    def row(name, px, bpp, fn, extra=None):
        ms = timed(fn, k5)
        out[name] = {"frames": n, "mpixel_per_s": px / ms / 1e3, "bytes_per_px": bpp, "gbs": px * bpp / ms / 1e6,
                     "frac": px * bpp / ms / 1e6 / peak}
        if extra:
            out[name].update(extra)

    for f, fname in ((1, "horizontal"), (2, "vertical"), (3, "gradient")):
        fj = (_abi.AlphaFilterJob * n)(*[_abi.AlphaFilterJob(src.ptr + i * st * H, dst.ptr + i * st * H, st, st, W, H, f, 0) for i in range(n)])
        row(f"alpha_unfilter_{fname}_1080p", n * W * H, 2.0, lambda: ctx._chk(ctx.lib.b2_alpha_unfilter(ctx.h, 0, fj, n)))
        row(f"alpha_filter_{fname}_1080p", n * W * H, 2.0, lambda: ctx._chk(ctx.lib.b2_alpha_filter(ctx.h, 0, fj, n)))
    dj = (_abi.AlphaDequantJob * n)(*[_abi.AlphaDequantJob(src.ptr + i * st * H, dst.ptr + i * st * H, st, st, W, H, 100, 0) for i in range(n)])
    row("alpha_level_smoothing_1080p", n * W * H, 3.0, lambda: ctx._chk(ctx.lib.b2_alpha_dequantize_levels(ctx.h, 0, dj, n, 1)))
    rj = (_abi.RescaleJob * n)(*[_abi.RescaleJob(src.ptr + i * st * H, dst.ptr + i * 4 * st * H, W, H, st, 2 * W, 2 * H, 2 * W, 1, 0) for i in range(n)])
    row("rescale_1080p_to_4k_luma", n * W * H, 5.0, lambda: ctx._chk(ctx.lib.b2_rescale_planes(ctx.h, 0, rj, n)), {"note": "Mpx/s of input; 1 B in + 4 B out per input pixel"})
    rj2 = (_abi.RescaleJob * n)(*[_abi.RescaleJob(src.ptr + i * st * H, dst.ptr + i * st * H, W, H, st, 1280, 720, 1280, 1, 0) for i in range(n)])
    row("rescale_1080p_to_720p_luma", n * W * H, 1.0 + 1280 * 720 / (W * H), lambda: ctx._chk(ctx.lib.b2_rescale_planes(ctx.h, 0, rj2, n)), {"note": "Mpx/s of input"})
    # the upscaled planes (dst: 128 x 4K) halved again: 4K -> 1080p, the 2 : 1 case of the shrinking kernel
    n4 = 64
    rj3 = (_abi.RescaleJob * n4)(*[_abi.RescaleJob(dst.ptr + i * 4 * st * H, src.ptr + i * st * H, 2 * W, 2 * H, 2 * st, W, H, st, 1, 0) for i in range(n4)])
    ms = timed(lambda: ctx._chk(ctx.lib.b2_rescale_planes(ctx.h, 0, rj3, n4)), k5)
    out["rescale_4k_to_1080p_luma"] = {"frames": n4, "mpixel_per_s": n4 * 4 * W * H / ms / 1e3, "bytes_per_px": 1.25, "gbs": n4 * 4 * W * H * 1.25 / ms / 1e6,
                                       "frac": n4 * 4 * W * H * 1.25 / ms / 1e6 / peak, "note": "Mpx/s of input; 1 B in + 0.25 B out per input pixel"}
    src.free()
    dst.free()
    return out


def cpu_baseline_block(args, alpha_mode):
    T = host_threads()
    n_cpu = int(min(args.frames, max(2 * T, 64)))
    # ~25 s of CPU work: passes over the sample for 1.5 s of wall time on all host threads
    rate, dt = cpu_import_rate(n_cpu, T, alpha_mode, min_seconds=max(1.5, 24.0 / T))
    rate1, _ = cpu_import_rate(4, 1, alpha_mode, min_seconds=1.0)
    srate, _ = cpu_import_rate(n_cpu, T, alpha_mode, scalar=True, min_seconds=1.0)
    srate1, _ = cpu_import_rate(4, 1, alpha_mode, scalar=True, min_seconds=1.0)
    return {
        "value": rate, "unit": "Mpixel/s", "cores": T, "kind": "port",
        "sample": f"{n_cpu} of {args.frames} frames, one image per task, repeated for {dt:.2f} s wall "
                  f"({dt * T:.0f} core-seconds) after one untimed pass",
        "single_thread": rate1, "libwebp160_simd_all_threads": libwebp_import_rate(n_cpu, T, alpha_mode),
        "scalar_no_vectorize_all_threads": srate, "scalar_no_vectorize_single_thread": srate1,
        "note": "oracle/oracle.c (gcc -O2) restating the reference; scalar_no_vectorize_* is the same source built "
                "-O2 -fno-tree-vectorize (the 'Go-equivalent scalar' figure of BASELINE.md 3); the Go reference does "
                "not compile and no Go toolchain exists here: GOMAXPROCS n/a"}


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores.
    The Go reference does not compile and no Go toolchain exists, so
    this is the oracle port, with all host threads, on the same config/metric."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    import oracle
    oracle.build()
    T = host_threads()
    alpha_mode = ALPHA_MODES[args.alpha]
    # bounded sample per step: sized from a probe so (K+W) steps take ~2 minutes at most
    probe_frames = max(T, 8)
    rate, dt = cpu_import_rate(probe_frames, T, alpha_mode)
    budget = 100.0 / max(1, args.steps + args.warmup)
    frames = int(min(args.frames, max(T, rate * 1e6 * min(budget, 4.0) / (W * H))))
    for _ in range(args.warmup):
        cpu_import_rate(frames, T, alpha_mode)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_import_rate(frames, T, alpha_mode)[1]
    value = frames * W * H * args.steps / t / 1e6
    sample = f"{frames} of {args.frames} 1080p frames per step, one image per task"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": T, "kind": "port", "sample": sample,
                         "note": "gcc -O2 C restatement of the reference (oracle/oracle.c); Go reference "
                                 "cannot be built: GOMAXPROCS n/a"},
        "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--alpha", default="sparse", choices=sorted(ALPHA_MODES))
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4],
                    help="BASELINE.json configs[k] as the headline of the line (default 1)")
    ap.add_argument("--no-extras", action="store_true", help="skip other kernels / other configs / cpu baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)  # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference(args)

    D = Dist()
    from go_webp_b200.device import Context
    peak, peak_src = hbm_peak()
    ctx = Context(D.local, n_streams=4)
    sampler = ClockSampler(D.local)
    sampler.start()
    line, pipe = config1_import(ctx, D, args, peak, peak_src, sampler)
    alpha_mode = ALPHA_MODES[args.alpha]
    if D.rank == 0 and D.world == 1 and not args.no_extras:
        line["cpu_baseline"] = cpu_baseline_block(args, alpha_mode)
        alpha_mixes = {}
        for name, am in ALPHA_MODES.items():
            pipe.fill_synthetic(SEED, first_img=0, alpha_mode=am)
            for _ in range(3):
                pipe.run_device()
            ctx.sync()
            ctx.event_record(4)
            for _ in range(10):
                pipe.run_device()
            ctx.event_record(5)
            ctx.sync()
            m = ctx.elapsed_ms(4, 5) / 10
            alpha_mixes[name] = {"mpixel_per_s": args.frames * W * H / m / 1e3,
                                 "frac": args.frames * W * H * IMPORT_BYTES_PER_PX / m / 1e6 / peak}
        line["alpha_mixes"] = alpha_mixes
    pipe.close()
    sampler.stop_flag = True
    if not args.no_extras:
        others = {}
        for name, fn in (("configs[3]", config3_upsample), ("configs[2]", config2_vp8l), ("configs[4]", config4_mixed)):
            try:
                others[name] = fn(ctx, D, peak, args.steps)
            except Exception as e:   # a failed side config must not take the headline down
                others[name] = {"error": repr(e)[:300]}
                D.barrier()
        line["other_configs"] = others
        k = f"configs[{args.config}]"
        if args.config != 1 and "error" not in others.get(k, {"error": 1}):
            o = others[k]
            line["headline_config"] = k
            line["configs[1]"] = {key: line[key] for key in ("value", "ms_per_step", "e2e", "roofline", "config", "scaling")}
            line.update({"value": o["value"], "ms_per_step": o["ms_per_step"], "scaling": o["scaling"],
                         "config": {"workload": o["workload"]}, "roofline": o["roofline"]})
            if "e2e" in o:
                line["e2e"] = o["e2e"]
        if D.rank == 0 and D.world == 1:
            line["other_kernels"] = other_kernels(ctx, peak, args.steps)
    if D.rank == 0:
        print(json.dumps(line))
    D.close()
    return 0


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the import kernel
    from the committed ncu capture (profiles/import_traffic.json), else None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "import_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        return None


if __name__ == "__main__":
    sys.exit(main())
