"""Hypothesis sharding across the GPUs of one node.

Every (pair, hypothesis) score is independent, so rank r of P scores the slice
[r*ceil(N/P), min(N,(r+1)*ceil(N/P))) of the rotation set for ALL pairs; the
only exchange is one small all-gather of the per-rank top-k lists
(B*k*12 bytes per rank) followed by a deterministic local merge, after which
every rank holds bit-identical results.  The reference has no multi-GPU
inference (it runs batch 1 on one GPU, test_co3d.py:133).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


def shard_bounds(N: int, rank: int, world: int) -> tuple[int, int]:
    per = -(-N // world)
    lo = min(N, rank * per)
    return lo, min(N, lo + per)


def all_gather_topk(val: torch.Tensor, idx: torch.Tensor, group=None):
    """[B,k] per rank -> [P,B,k] on every rank (one all-gather each for the
    fp32 values and the int64 indices)."""
    world = dist.get_world_size(group)
    val, idx = val.contiguous(), idx.contiguous()
    vals = [torch.empty_like(val) for _ in range(world)]
    idxs = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(vals, val, group=group)
    dist.all_gather(idxs, idx, group=group)
    vals, idxs = torch.stack(vals), torch.stack(idxs)
    return vals, idxs


def pad_topk(val: torch.Tensor, idx: torch.Tensor, k: int):
    """Ranks whose slice is shorter than k pad with (-inf, -1), which the merge ignores."""
    B, have = val.shape
    if have == k:
        return val, idx
    pv = torch.full((B, k - have), float("-inf"), dtype=val.dtype, device=val.device)
    pi = torch.full((B, k - have), -1, dtype=idx.dtype, device=idx.device)
    return torch.cat([val, pv], 1), torch.cat([idx, pi], 1)


class ShardedVerifier:
    """Wraps a HypothesisVerifier; `score` takes the FULL rotation set on every
    rank (36 B per hypothesis, replicated like the 40 KB/pair volumes) and
    scores only this rank's slice."""

    def __init__(self, verifier, group=None, merge=None, score_fn=None):
        self.verifier = verifier
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._merge = merge or ops.topk_merge
        # score_fn(vol_src, vol_tgt, R_slice, k, idx_offset) -> (val [B,k'], idx [B,k'])
        self._score_fn = score_fn or self._score_local

    def _score_local(self, vol_src, vol_tgt, R_slice, k, idx_offset):
        r = self.verifier.score(vol_src, vol_tgt, R_slice, k=k, return_scores=False, idx_offset=idx_offset, gather=False)
        return r.topk_val, r.topk_idx

    @torch.no_grad()
    def score(self, vol_src, vol_tgt, R, k: int = 1):
        """Returns (topk_val [B,k], topk_idx [B,k] global, R_best [B,k,3,3])."""
        per_pair = R.dim() == 4
        N = R.shape[1] if per_pair else R.shape[0]
        k = min(k, N)
        lo, hi = shard_bounds(N, self.rank, self.world)
        B = vol_src.shape[0]
        if hi > lo:
            Rs = (R[:, lo:hi] if per_pair else R[lo:hi]).contiguous()
            val, idx = self._score_fn(vol_src, vol_tgt, Rs, min(k, hi - lo), lo)
        else:
            val = torch.empty(B, 0, dtype=torch.float32, device=vol_src.device)
            idx = torch.empty(B, 0, dtype=torch.int64, device=vol_src.device)
        val, idx = pad_topk(val, idx, k)
        if self.world > 1:
            vals, idxs = all_gather_topk(val, idx, self.group)
            val, idx = self._merge(vals, idxs)
        safe = idx.clamp_min(0)
        if per_pair:
            R_best = torch.gather(R, 1, safe[..., None, None].expand(-1, -1, 3, 3))
        else:
            R_best = R[safe]
        return val, idx, R_best
