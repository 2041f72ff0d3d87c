"""Time-sharding of a sliding-window search across the GPUs of one box.

The reference is single-process (MLGWSC-1/inference.py:532-589 loops over segments, :465-487 over
256-window batches).  Windows are independent, so the only exchange in the multi-GPU path is the
gather of the small per-rank trigger lists at the end; there is no data-path collective.

Design
  * unit of work = one 256-window batch of one segment (the reference's DataLoader batch,
    inference.py:465).  Shards are made of WHOLE batches so that batch-coupled front ends (QScan
    picks one Q plane per call from the global max over the batch) see exactly the
    batches a single-GPU run sees.
  * batches are assigned to ranks as contiguous runs (balanced prefix sums), so a rank reads one
    contiguous sample range per segment plus a (2048 - hop)-sample halo.
  * trigger exchange: all_gather of counts, then all_gather of lists padded to the max count
    (int64 window index, f32 score), then a stable sort by (segment, window) on every rank.
    Works with backend "nccl" (CUDA tensors) and "gloo" (CPU tensors; used by the tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

BATCH = 256
WINDOW = 2048


@dataclass(frozen=True)
class ShardPiece:
    """A contiguous run of windows of one segment owned by one rank."""
    segment: int          # index into the caller's segment list
    first_window: int     # first window (multiple of BATCH)
    n_windows: int

    def sample_range(self, hop: int, window: int = WINDOW) -> Tuple[int, int]:
        """[lo, hi) sample range of the segment this piece reads (includes the halo)."""
        lo = self.first_window * hop
        return lo, lo + (self.n_windows - 1) * hop + window


def n_windows(n_samples: int, hop: int, window: int = WINDOW) -> int:
    """len(SegmentSlicer) (inference.py:247-252)."""
    return 0 if n_samples < window else 1 + (n_samples - window) // hop


def plan_shards(windows_per_segment: Sequence[int], world: int, batch: int = BATCH) -> List[List[ShardPiece]]:
    """Assign whole batches to `world` ranks as contiguous, balanced runs.  Returns, per rank, the
    list of pieces in (segment, window) order.  Deterministic; every window is covered once."""
    if world < 1:
        raise ValueError("world must be >= 1")
    batches: List[Tuple[int, int, int]] = []   # (segment, first_window, n)
    for s, nw in enumerate(windows_per_segment):
        for k0 in range(0, int(nw), batch):
            batches.append((s, k0, min(batch, int(nw) - k0)))
    total = sum(b[2] for b in batches)
    plan: List[List[ShardPiece]] = [[] for _ in range(world)]
    if total == 0:
        return plan
    # rank r owns the batches whose cumulative window offset falls in [r, r+1) * total / world
    done = 0
    for seg, k0, n in batches:
        r = min(world - 1, (done * world) // total)
        pieces = plan[r]
        if pieces and pieces[-1].segment == seg and pieces[-1].first_window + pieces[-1].n_windows == k0:
            last = pieces[-1]
            pieces[-1] = ShardPiece(seg, last.first_window, last.n_windows + n)
        else:
            pieces.append(ShardPiece(seg, k0, n))
        done += n
    return plan


def gather_triggers(seg: torch.Tensor, idx: torch.Tensor, score: torch.Tensor,
                    group: Optional[dist.ProcessGroup] = None
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """All-gather per-rank trigger lists (segment id, window index, score) and return the merged
    list sorted by (segment, window) -- identical on every rank.  Without an initialised process
    group this is the identity (single-GPU run)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        order = torch.argsort(seg * (1 << 40) + idx, stable=True)
        return seg[order], idx[order], score[order]
    world = dist.get_world_size(group)
    dev = idx.device
    cnt = torch.tensor([idx.numel()], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    counts = [int(c.item()) for c in cnts]
    cap = max(max(counts), 1)
    keys = torch.full((cap, 2), -1, dtype=torch.int64, device=dev)
    vals = torch.zeros(cap, dtype=torch.float32, device=dev)
    n = idx.numel()
    keys[:n, 0], keys[:n, 1], vals[:n] = seg.to(torch.int64), idx.to(torch.int64), score.to(torch.float32)
    all_keys = [torch.empty_like(keys) for _ in range(world)]
    all_vals = [torch.empty_like(vals) for _ in range(world)]
    dist.all_gather(all_keys, keys, group=group)
    dist.all_gather(all_vals, vals, group=group)
    k = torch.cat([a[:c] for a, c in zip(all_keys, counts)])
    v = torch.cat([a[:c] for a, c in zip(all_vals, counts)])
    order = torch.argsort(k[:, 0] * (1 << 40) + k[:, 1], stable=True)
    return k[order, 0], k[order, 1], v[order]


def gather_scores(pieces: Sequence[ShardPiece], scores: Sequence[torch.Tensor],
                  windows_per_segment: Sequence[int], group: Optional[dist.ProcessGroup] = None,
                  device: Optional[torch.device] = None) -> List[torch.Tensor]:
    """Reassemble the per-window scores (`all_vals` of the reference) of every segment on every
    rank: each rank contributes its pieces; result[s] is a [windows_per_segment[s]] f32 tensor."""
    dev = device if device is not None else (scores[0].device if len(scores) else torch.device("cpu"))
    total = int(sum(windows_per_segment))
    offs = [0]
    for nw in windows_per_segment:
        offs.append(offs[-1] + int(nw))
    flat = torch.zeros(total, dtype=torch.float32, device=dev)
    for p, sc in zip(pieces, scores):
        flat[offs[p.segment] + p.first_window: offs[p.segment] + p.first_window + p.n_windows] = sc
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        # pieces are disjoint, untouched entries are zero: a sum is a gather
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return [flat[offs[s]:offs[s + 1]] for s in range(len(windows_per_segment))]


def sharded_search(network, segments: Sequence[torch.Tensor], hop: int, threshold: float,
                   rank: int = 0, world: int = 1, group: Optional[dist.ProcessGroup] = None,
                   batch: int = BATCH, device: Optional[torch.device] = None):
    """Search `segments` (list of [D, n_samples] strain tensors; host or device resident) with
    `network.stream_search(strain, hop, n_windows, thr, first_window)`.  Every rank processes its own pieces --
    it touches (and, for host-resident segments, uploads) only the sample range of each piece plus the
    (2048 - hop)-sample halo, `ShardPiece.sample_range` -- and all ranks return the same (seg, window, score)
    trigger list and the per-segment score arrays.  Window indices are global (segment-relative)."""
    nws = [n_windows(int(s.shape[-1]), hop) for s in segments]
    plan = plan_shards(nws, world, batch)
    mine = plan[rank]
    t_seg, t_idx, t_sc, sc_list = [], [], [], []
    dev = device
    for p in mine:
        lo, hi = p.sample_range(hop)
        strain = segments[p.segment][:, lo:hi]
        if device is not None and strain.device != torch.device(device):
            strain = strain.to(device, non_blocking=True)
        strain = strain.contiguous()
        # the slice starts at the piece's first window: local window k == global window first_window + k
        scores, idx, sc = network.stream_search(strain, hop, p.n_windows, threshold, first_window=0)
        dev = scores.device
        sc_list.append(scores)
        t_idx.append(idx.to(torch.int64) + p.first_window)
        t_sc.append(sc.to(torch.float32))
        t_seg.append(torch.full_like(idx, p.segment, dtype=torch.int64))
    if dev is None:
        dev = segments[0].device if len(segments) else torch.device("cpu")
    cat = (lambda xs, dt: torch.cat(xs) if xs else torch.empty(0, dtype=dt, device=dev))
    seg, idx, sc = gather_triggers(cat(t_seg, torch.int64), cat(t_idx, torch.int64), cat(t_sc, torch.float32), group)
    if not sc_list:
        sc_list, mine = [torch.empty(0, device=dev)], []
    all_scores = gather_scores(mine, sc_list, nws, group, device=dev) if nws else []
    return (seg, idx, sc), all_scores
