#!/usr/bin/env python
"""configs[4]: mixed-size batch (64x64 .. 8192x8192, odd widths/heights) through the full
pixel pipeline (import -> fancy upsample -> VP8L inverse chain), LPT-partitioned by
algorithmic bytes over G GPUs.

  python tools/mixed_bench.py                 one GPU: times every rank's share of the
                                              G = 1, 2, 4, 8 partitions one after the other
                                              (job time at G GPUs = the slowest share)
  torchrun --nproc-per-node G tools/mixed_bench.py
                                              G GPUs: rank r runs its share, max over ranks
Prints one JSON line per G."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from go_webp_b200 import _abi  # noqa: E402
if os.environ.get("B2_LIB"):  # experiment builds (tools/variants.sh)
    _abi.LIB_PATH = os.environ["B2_LIB"]
from go_webp_b200.device import Context  # noqa: E402
from go_webp_b200.partition import import_cost_bytes, lpt_partition, upsample_cost_bytes  # noqa: E402
from go_webp_b200.pipeline import MixedPipeline, mixed_sizes  # noqa: E402

N_IMAGES = int(os.environ.get("B2_MIXED_IMAGES", "256"))
SEED = 0xB2000005
PEAK = 6552.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def cost(w, h):
    return import_cost_bytes(w, h) + upsample_cost_bytes(w, h, True) + 8 * w * h


def time_share(ctx, sizes, steps=5, warmup=2):
    p = MixedPipeline(ctx, sizes)
    p.fill_synthetic(SEED)
    for _ in range(warmup):
        p.run_device()
    ctx.sync()
    ctx.event_record(0)
    for _ in range(steps):
        p.run_device()
    ctx.event_record(1)
    ctx.sync()
    ms = ctx.elapsed_ms(0, 1) / steps
    px, by = p.pixels, p.algorithmic_bytes()
    p.close()
    return ms, px, by


def breakdown(ctx, sizes, label):
    """Per-kernel device times of one pass over `sizes` (B2_MIXED_BREAKDOWN=1)."""
    from go_webp_b200 import _abi
    p = MixedPipeline(ctx, sizes)
    p.fill_synthetic(SEED)
    c = ctx
    calls = [("import", lambda: c._chk(c.lib.b2_rgba_to_yuv420a(c.h, 0, p._imp, p.n, _abi.ORDER_RGBA, _abi.IMPORT_ALPHA, p.d_flags.ptr)), 6.5),
             ("upsample", lambda: c._chk(c.lib.b2_yuv420a_to_rgba_fancy(c.h, 0, p._ups, p.n, p.mode)), 6.5),
             ("chain", lambda: c._chk(c.lib.b2_vp8l_inverse_chain(c.h, 0, p._arg, p.n)), 8.0)]
    out = {"subset": label, "n_images": len(sizes), "mpixel": p.pixels / 1e6}
    for name, fn, bpp in calls:
        for _ in range(2):
            fn()
        c.sync(); c.event_record(0)
        for _ in range(5):
            fn()
        c.event_record(1); c.sync()
        ms = c.elapsed_ms(0, 1) / 5
        out[name] = {"ms": ms, "mpixel_per_s": p.pixels / ms / 1e3, "frac": p.pixels * bpp / ms / 1e6 / PEAK}
    p.close()
    print(json.dumps(out), flush=True)


def main():
    sizes = mixed_sizes(N_IMAGES, SEED)
    if os.environ.get("B2_MIXED_BREAKDOWN"):
        ctx = Context(0, 1)
        breakdown(ctx, sizes, "all")
        breakdown(ctx, [s for s in sizes if s[0] % 4 == 0], "width % 4 == 0")
        breakdown(ctx, [s for s in sizes if s[0] % 2 == 1], "odd width")
        breakdown(ctx, [s for s in sizes if s[0] % 4 == 2], "width % 4 == 2")
        return
    costs = [cost(w, h) for (w, h) in sizes]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl")
        ctx = Context(local, 1)
        mine = lpt_partition(costs, world)[rank]
        ms, px, by = time_share(ctx, [sizes[i] for i in mine])
        t = torch.tensor([ms, float(px), float(by)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        if rank == 0:
            print(json.dumps({"workload": "configs[4] mixed sizes, full pipeline", "n_images": N_IMAGES, "n_gpus": world,
                              "mode": "one rank per GPU", "ms_per_pass": tmax[0].item(), "mean_ms": tsum[0].item() / world,
                              "mpixel_per_s": tsum[1].item() / tmax[0].item() / 1e3,
                              "gbs_per_gpu": tsum[2].item() / world / tmax[0].item() / 1e6,
                              "imbalance_max_over_mean": tmax[0].item() * world / tsum[0].item()}))
        dist.destroy_process_group()
        return
    ctx = Context(0, 1)
    if os.environ.get("B2_MIXED_CALIBRATED"):   # the time-based cost model bench.py uses
        from go_webp_b200.partition import calibrated_costs
        from go_webp_b200.pipeline import calibrate_mixed_rates
        rates = calibrate_mixed_rates(ctx, sizes)
        print(json.dumps({"rates_GBps": {k: round(v / 1e6, 1) for k, v in rates.items()}}), flush=True)
        costs = calibrated_costs(sizes, rates)
    for g in (1, 2, 4, 8):
        parts = lpt_partition(costs, g)
        res = [time_share(ctx, [sizes[i] for i in part]) for part in parts]
        if os.environ.get("B2_MIXED_REFINE") and g > 1:   # measured shares correct the model, a few rounds
            cur = list(costs)
            best = (max(r[0] for r in res), parts, res)
            for it in range(int(os.environ["B2_MIXED_REFINE"])):
                pred = [sum(cur[i] for i in part) for part in parts]
                meas = [r[0] for r in res]
                norm = sum(meas) / sum(pred)
                for part, pr, me in zip(parts, pred, meas):
                    f = (me / (pr * norm)) ** 0.7
                    for i in part:
                        cur[i] *= f
                parts = lpt_partition(cur, g)
                res = [time_share(ctx, [sizes[i] for i in part]) for part in parts]
                tm = max(r[0] for r in res)
                print(json.dumps({"refine_round": it + 1, "n_gpus": g, "imbalance": tm * g / sum(r[0] for r in res)}), flush=True)
                if tm < best[0]:
                    best = (tm, parts, res)
            _, parts, res = best
        tmax = max(r[0] for r in res)
        tmean = sum(r[0] for r in res) / g
        px = sum(r[1] for r in res)
        by = sum(r[2] for r in res)
        print(json.dumps({"workload": "configs[4] mixed sizes, full pipeline", "n_images": N_IMAGES, "n_gpus": g,
                          "mode": "shares timed one after the other on one GPU", "ms_per_pass": tmax, "mean_ms": tmean,
                          "mpixel_per_s": px / tmax / 1e3, "gbs_per_gpu": by / g / tmax / 1e6,
                          "frac_of_measured_peak_per_gpu": by / g / tmax / 1e6 / PEAK,
                          "imbalance_max_over_mean": tmax / tmean,
                          "cost_imbalance": max(sum(costs[i] for i in part) for part in parts) * g / sum(costs)}), flush=True)


if __name__ == "__main__":
    main()
