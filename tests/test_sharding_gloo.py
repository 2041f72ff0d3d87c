"""Multi-GPU host logic on CPU: shard planning + trigger/score gather with the gloo backend,
world_size 2 and 3.  The per-window network is a deterministic stand-in
(the GPU kernels are covered by the -m gpu tests); what is checked here is that a sharded search
returns exactly what the single-process search returns."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gw_whisper_b200 import sharding as S

HOP = 204


class FakeNetwork:
    """score(window k of a segment) = mean of the window's first detector: cheap, exact, depends on
    the samples the shard must own (halo included)."""

    def stream_search(self, strain, hop, n_windows, thr, first_window=0):
        # a shard is handed exactly its sample range plus the halo, never the whole segment
        assert first_window == 0 and strain.shape[1] == (n_windows - 1) * hop + 2048
        k = torch.arange(first_window, first_window + n_windows)
        idx = k[:, None] * hop + torch.arange(2048)[None, :]
        scores = strain[0][idx].double().mean(dim=1).float()
        keep = (scores > thr).nonzero().flatten()
        return scores, k[keep], scores[keep]


def _segments():
    g = torch.Generator().manual_seed(99)
    lens = [2048 + HOP * 700, 2048 + HOP * 255, 2048 + HOP * 256, 4000, 1000]  # incl. ragged and too-short
    return [torch.randn(2, n, generator=g) for n in lens]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        (seg, idx, sc), scores = S.sharded_search(FakeNetwork(), _segments(), HOP, 0.01, rank, world)
        torch.save({"seg": seg, "idx": idx, "sc": sc, "scores": scores}, os.path.join(out, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_plan_covers_every_window_once_and_keeps_batches_whole():
    nws = [701, 256, 257, 10, 0]
    for world in (1, 2, 3, 8):
        plan = S.plan_shards(nws, world)
        assert len(plan) == world
        seen = [np.zeros(n, dtype=int) for n in nws]
        for pieces in plan:
            for p in pieces:
                assert p.first_window % S.BATCH == 0          # shards start on batch boundaries
                assert p.n_windows % S.BATCH == 0 or p.first_window + p.n_windows == nws[p.segment]
                seen[p.segment][p.first_window:p.first_window + p.n_windows] += 1
        assert all((s == 1).all() for s in seen)
        loads = [sum(p.n_windows for p in pieces) for pieces in plan]
        assert max(loads) - min(loads) <= 2 * S.BATCH or world > 5


def test_plan_empty_and_degenerate():
    assert S.plan_shards([], 4) == [[], [], [], []]
    assert S.plan_shards([0, 0], 2) == [[], []]
    assert S.n_windows(2047, HOP) == 0 and S.n_windows(2048, HOP) == 1 and S.n_windows(2048 + 203, HOP) == 1
    assert S.n_windows(2048 + 204, HOP) == 2
    with pytest.raises(ValueError):
        S.plan_shards([5], 0)


def test_sample_range_includes_halo():
    p = S.ShardPiece(0, 256, 512)
    lo, hi = p.sample_range(HOP)
    assert lo == 256 * HOP and hi == (256 + 511) * HOP + 2048


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_matches_single_process(tmp_path, world):
    (seg0, idx0, sc0), scores0 = S.sharded_search(FakeNetwork(), _segments(), HOP, 0.01, 0, 1)
    assert seg0.numel() > 10
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert torch.equal(got["seg"], seg0) and torch.equal(got["idx"], idx0)
        assert torch.equal(got["sc"], sc0)                       # bit-exact: same windows, same arithmetic
        assert len(got["scores"]) == len(scores0)
        for a, b in zip(got["scores"], scores0):
            assert torch.equal(a, b)
