"""world_size-2 gloo test of the N>1 path's host logic (no GPU): every rank
derives its shard alone (no data-path collective), shards are disjoint and
cover the batch, and the max-over-ranks timing reduction bench.py uses works."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from go_webp_b200 import partition
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 37
    mine = list(partition.contiguous_shard(n, world, rank))
    costs = [(i * 7919) % 1000 + 1 for i in range(n)]
    lpt = partition.lpt_partition(costs, world)[rank]
    # verification-only collectives (the data path itself has none)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, lpt))
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((gathered, float(t)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cont = sorted(i for g in gathered for i in g[0])
    lpt = sorted(i for g in gathered for i in g[1])
    assert cont == list(range(37)) and lpt == list(range(37))
    assert tmax == 2.0


def test_mixed_size_batch_and_lpt_balance():
    """configs[4] batch generator: deterministic, corner cases first, the
    four parity classes equally represented, sizes within [64, 8192]; LPT on algorithmic
    bytes balances 2/4/8 ranks to well under 1 %."""
    sys.path.insert(0, ROOT)
    from go_webp_b200 import partition
    from go_webp_b200.pipeline import mixed_sizes
    s = mixed_sizes(256)
    assert s == mixed_sizes(256)
    assert s[:4] == [(64, 64), (8192, 8192), (65, 8191), (8191, 65)]
    assert all(64 <= w <= 8192 and 64 <= h <= 8192 for (w, h) in s)
    classes = [0, 0, 0, 0]
    for (w, h) in s[4:]:
        classes[(w & 1) | ((h & 1) << 1)] += 1
    assert max(classes) - min(classes) <= 1
    assert 0.4e9 < sum(w * h for (w, h) in s) < 1.2e9
    costs = [partition.import_cost_bytes(w, h) + partition.upsample_cost_bytes(w, h, True) + 8 * w * h for (w, h) in s]
    for g in (2, 4, 8):
        parts = partition.lpt_partition(costs, g)
        assert sorted(i for p in parts for i in p) == list(range(256))
        assert max(sum(costs[i] for i in p) for p in parts) * g / sum(costs) < 1.01


def test_calibrated_cost_model_balances_time_not_bytes():
    """configs[4], round 2: LPT on bytes / rate(class).  With rates that differ by class the time
    imbalance of the byte model shows, the calibrated model removes it; missing classes fall back."""
    sys.path.insert(0, ROOT)
    from go_webp_b200 import partition
    from go_webp_b200.pipeline import mixed_sizes
    s = mixed_sizes(256)
    classes = sorted({partition.size_class(w, h) for (w, h) in s})
    assert len(classes) >= 8 and all(c.split("/")[0] in ("w16", "w4", "w2", "w1") for c in classes)
    # a plausible machine: odd widths and narrow rows run slower
    true_rate = {c: (1.0 if c.startswith("w16") else 0.8 if c.startswith("w4") else 0.55) *
                 (0.5 if c.endswith("narrow") else 1.0) for c in classes}
    t_true = partition.calibrated_costs(s, true_rate)
    byte_costs = [partition.full_pipeline_cost_bytes(w, h) for (w, h) in s]
    for g in (2, 4, 8):
        by_bytes = partition.lpt_partition(byte_costs, g)
        by_time = partition.lpt_partition(t_true, g)
        imb = lambda parts: max(sum(t_true[i] for i in p) for p in parts) * g / sum(t_true)  # noqa: E731
        assert imb(by_time) < 1.01
        assert imb(by_time) <= imb(by_bytes)
    # unmeasured classes use the mean rate: costs stay finite and proportional to bytes
    c = partition.calibrated_costs([(64, 64), (1920, 1080)], {})
    assert c[1] / c[0] == pytest.approx(partition.full_pipeline_cost_bytes(1920, 1080) / partition.full_pipeline_cost_bytes(64, 64))
