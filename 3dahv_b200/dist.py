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


class PeerExchange:
    """Exchange buffers for the fused sharded step (`ahv_verify_sharded`): every rank allocates one buffer
    (`ahv_peer_alloc`), sends its CUDA IPC handle to the peers (one all-gather of 64 bytes at set-up) and maps
    theirs, after which the scoring kernels talk to each other through NVLink peer memory and a verification
    step involves no NCCL call at all.  One instance per process group; it serves any call with B <= max_pairs and
    k <= max_k (the entry stride depends on the capacity only, so B and k may change from step to step)."""

    def __init__(self, max_pairs: int, device, group=None, max_k: int = 1):
        import ctypes

        from . import _lib

        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError("the peer exchange covers the GPUs of one node (<= 8)")
        if not (1 <= max_k <= 32) or max_pairs < 1:
            raise ValueError("max_pairs >= 1 and 1 <= max_k <= 32 expected")
        self.max_pairs, self.max_k = max_pairs, max_k
        self.device = torch.device(device)
        lib = _lib.lib()
        self._lib = lib
        self.own, self.ptrs, self._opened = None, [], []
        # Every step below that can fail is local; the ranks agree on the outcome with one all-reduce at the
        # end, so a rank that cannot export or map a buffer makes ALL ranks raise instead of leaving the
        # others blocked in a collective.
        ok = True
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(self.device):
            if lib.ahv_peer_alloc(max_pairs, max_k, ctypes.byref(own)) != 0:
                ok = False
            elif lib.ahv_peer_export(own, handle) != 0:
                ok = False
            self.own = own.value
            mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(self.device)
            every = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(every, mine, group=group)
            flag = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag) == 1:
                for r, h in enumerate(every):
                    if r == self.rank:
                        self.ptrs.append(self.own)
                        continue
                    p = ctypes.c_void_p()
                    if lib.ahv_peer_open(bytes(h.cpu().numpy().tobytes()), ctypes.byref(p)) != 0:
                        ok = False
                        break
                    self.ptrs.append(p.value)
                    self._opened.append(p.value)
            else:
                ok = False
            flag = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag) != 1:
            self.close()
            raise RuntimeError("peer exchange unavailable: a rank could not allocate, export or map a CUDA IPC buffer "
                               "(use the NCCL path: ShardedVerifier without `peer`)")

    def check(self) -> int:
        """Synchronises the device and raises if an exchange on this rank timed out waiting for a peer (that
        step returned NaN scores and index -1).  Returns the number of exchanges completed."""
        import ctypes

        from . import _lib

        seq, err = ctypes.c_uint32(), ctypes.c_uint32()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.ahv_peer_status(self.own, ctypes.byref(seq), ctypes.byref(err)), "ahv_peer_status")
        if err.value != 0:
            raise RuntimeError("a sharded verification step timed out waiting for a peer rank")
        return int(seq.value)

    def close(self):
        """Collective: every rank unmaps the peers' buffers before anyone frees its own."""
        if getattr(self, "own", None) is None and not getattr(self, "_opened", None):
            return
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            for p in self._opened:
                self._lib.ahv_peer_close(p)
            if dist.is_initialized():
                dist.barrier(group=self.group)
            if self.own:
                self._lib.ahv_peer_free(self.own)
        self.own, self.ptrs, self._opened = None, [], []


class ShardedVerifier:
    """Wraps a HypothesisVerifier; `score` takes the FULL rotation set on every
    rank (36 B per hypothesis, replicated like the 40 KB/pair volumes) and
    scores only this rank's slice."""

    def __init__(self, verifier, group=None, merge=None, score_fn=None, peer: "PeerExchange | None" = None,
                 check_every: int = 0):
        self.verifier = verifier
        self.group = group
        self.peer = peer   # NVLink peer exchange by the kernels themselves (None: NCCL all-gather + merge kernel)
        self.check_every = check_every   # synchronising error check (a peer that never arrived) every so many steps
        self._steps = 0
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
        if (self.peer is not None and self.world > 1 and B <= self.peer.max_pairs and k <= self.peer.max_k
                and self._score_fn == self._score_local):
            # one C call: shard scoring + NVLink exchange + merge (k == 1: all inside the scoring kernel); empty
            # slices take part in the exchange with empty lists
            v = self.verifier
            W1, W2, b2 = v._weights_on(vol_src.device)
            Rs = (R[:, lo:hi] if per_pair else R[lo:hi]).contiguous()
            vs = vol_src if vol_src.dtype == torch.bfloat16 else vol_src.float()
            val, idx, Rb = ops.verify_sharded(vs, vol_tgt.float(), Rs, W1, W2, b2, lo, self.peer, k=k, math=v.math,
                                              workspace=v._workspace(B, hi - lo, k, vol_src.device))
            self._steps += 1
            if self.check_every and self._steps % self.check_every == 0:
                self.peer.check()
            return val, idx, Rb
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

    @torch.no_grad()
    def refine(self, vol_src, vol_tgt, R, k: int = 32, m: int = 64, max_angle_deg: float = 5.0, seed: int = 0):
        """Two-pass selection (BASELINE config 4) with BOTH passes sharded: pass 1 scores this rank's slice of the
        dense set and the ranks exchange their top-k (identical on every rank afterwards, rotations included);
        every rank then builds the same k*m candidates per pair (`ahv_so3_perturb`) and pass 2 shards THOSE over the
        ranks, winners exchanged inside the scoring kernel.  With a `peer` exchange the whole thing contains no NCCL
        call.  Returns (R_best [B,3,3], score [B], (first_val, first_idx, first_R), candidates [B,k*m,3,3]) - equal
        to `HypothesisVerifier.refine` on one GPU."""
        from . import so3

        per_pair = R.dim() == 4
        k = min(k, R.shape[1] if per_pair else R.shape[0])
        first_val, first_idx, first_R = self.score(vol_src, vol_tgt, R, k=k)
        B = vol_src.shape[0]
        cand = so3.perturb_rotations(first_R.contiguous(), m, max_angle_deg, seed).reshape(B, k * m, 3, 3).contiguous()
        val, _, R_best = self.score(vol_src, vol_tgt, cand, k=1)
        return R_best[:, 0], val[:, 0], (first_val, first_idx, first_R), cand

    @torch.no_grad()
    def scores(self, vol_src, vol_tgt, R):
        """The full pred_sim [B,N] (modules/model.py:193) of a sharded set on every rank: each rank scores its slice,
        one NCCL all-gather of B*N/P floats per rank puts the matrix together (25.6 MB per rank at
        B=128, N=50 000, P=8).  Only needed when the caller wants every score; selection never moves them."""
        per_pair = R.dim() == 4
        N = R.shape[1] if per_pair else R.shape[0]
        B = vol_src.shape[0]
        per = -(-N // self.world)
        lo, hi = shard_bounds(N, self.rank, self.world)
        part = torch.zeros(B, per, device=vol_src.device, dtype=torch.float32)     # equal-sized pieces for the all-gather
        if hi > lo:
            Rs = (R[:, lo:hi] if per_pair else R[lo:hi]).contiguous()
            part[:, :hi - lo] = self.verifier.score(vol_src, vol_tgt, Rs, k=1, return_scores=True).scores
        if self.world == 1:
            return part[:, :N]
        pieces = [torch.empty_like(part) for _ in range(self.world)]
        dist.all_gather(pieces, part, group=self.group)
        return torch.cat(pieces, dim=1)[:, :N]
