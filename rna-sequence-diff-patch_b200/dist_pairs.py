"""dist_pairs.py — a pair batch sharded over the GPUs of one box.

Pairs are independent units: the pair list is cut into contiguous ranges balanced by the number of
matrix cells (sum of m*n, not the pair count), every rank scores its own range, and the results are
concatenated in pair order.  There is no exchange on the data path; the only communication is the
final gather of 8 bytes per pair.  torch is plumbing only (process group)."""
from __future__ import annotations

import numpy as np

from .dist_search import slice_packed
from .encoding import PackedSeqs


def pair_shard_bounds(a_len: np.ndarray, b_len: np.ndarray, world: int):
    """Contiguous [lo, hi) pair ranges per rank, balanced by sum(m*n) (8e: 'balance by Σ m·n, not by count')."""
    n = int(a_len.shape[0])
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(world, 1) - 1)
    cells = a_len.astype(np.int64) * b_len.astype(np.int64) + 1      # +1: empty pairs still cost a launch slot
    csum = np.concatenate([[0], np.cumsum(cells)])
    total = int(csum[-1])
    cuts = [0] + [int(np.searchsorted(csum, total * r // world, side="left")) for r in range(1, world)] + [n]
    for r in range(1, len(cuts)):
        cuts[r] = min(max(cuts[r], cuts[r - 1]), n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_concat(local: np.ndarray, bounds, group=None, device=None) -> np.ndarray:
    """Every rank contributes result[lo:hi]; returns the whole array on every rank (NCCL on `device`, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    width = max(hi - lo for lo, hi in bounds)
    pad = np.zeros(width, local.dtype); pad[:local.shape[0]] = local
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * width, dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    out = out.cpu().numpy().reshape(world, width)
    return np.concatenate([out[r, :hi - lo] for r, (lo, hi) in enumerate(bounds)])


class ShardedPairs:
    """distance_batch / script_batch of one rank's share of a pair list."""

    def __init__(self, engine, rank: int = 0, world: int = 1, group=None, device=None):
        self.engine, self.rank, self.world, self.group, self.device = engine, rank, world, group, device

    def local_range(self, A: PackedSeqs, B: PackedSeqs):
        bounds = pair_shard_bounds(A.len, B.len, self.world)
        return bounds, bounds[self.rank]

    def distance_batch(self, A: PackedSeqs, B: PackedSeqs, force_mode: int = 0) -> np.ndarray:
        """All ranks pass the same (A, B); every rank gets all distances back, in pair order."""
        bounds, (lo, hi) = self.local_range(A, B)
        a, b = slice_packed(A, lo, hi), slice_packed(B, lo, hi)
        a.symmask, b.symmask = A.symmask, B.symmask              # one numeric mode for the whole list
        local = self.engine.distance_batch(a, b, force_mode=force_mode) if hi > lo else np.zeros(0, np.float64)
        return gather_concat(local, bounds, self.group, self.device)

    def script_batch_local(self, A: PackedSeqs, B: PackedSeqs, **kw):
        """Edit scripts of this rank's range only (scripts stay where they were produced): -> (lo, hi, scripts)."""
        _, (lo, hi) = self.local_range(A, B)
        a, b = slice_packed(A, lo, hi), slice_packed(B, lo, hi)
        a.symmask, b.symmask = A.symmask, B.symmask
        return lo, hi, self.engine.script_batch(a, b, **kw)
