"""HypothesisVerifier — the fused hypothesis-and-verification step.

Replaces the idiom the reference pastes at modules/model.py:131-146, :184-196,
test_co3d.py:137-146 and test_linemod.py:43-63:

    rotate_volume(vol.expand(N), R) -> forward_3d2d -> (a*b[:,None]).sum(2).mean(-1)
    -> torch.max(dim=1) -> R[idx]

with one call into lib3dahv_b200 (no rotated volume, grid or feature tensor is
ever materialised).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import ops
from ._lib import MATH_FP32, MATH_TC, MATH_TC_F16GATHER

DEFAULT_MATH = MATH_TC


@dataclass
class VerifyResult:
    scores: torch.Tensor | None    # [B,N] pred_sim (modules/model.py:193)
    topk_val: torch.Tensor         # [B,k]
    topk_idx: torch.Tensor         # [B,k] int64, global hypothesis index
    R_best: torch.Tensor           # [B,k,3,3] sampled_R[pred_index] (:196)


class HypothesisVerifier:
    """Holds the verification-head weights (state-dict names
    `feature_aligner.feature_embedding_2d.{0.weight,2.weight,2.bias}`)."""

    def __init__(self, W1: torch.Tensor, W2: torch.Tensor, b2: torch.Tensor, math: int | None = None):
        self.W1 = W1.detach().reshape(32, 384).float().contiguous()
        self.W2 = W2.detach().reshape(32, 32).float().contiguous()
        self.b2 = b2.detach().float().contiguous()
        self.math = DEFAULT_MATH if math is None else math
        self._ws = None

    @classmethod
    def from_feature_aligner(cls, feature_aligner, math: int | None = None) -> "HypothesisVerifier":
        head = feature_aligner.feature_embedding_2d
        return cls(head[0].weight, head[2].weight, head[2].bias, math)

    def to(self, device) -> "HypothesisVerifier":
        self.W1, self.W2, self.b2 = self.W1.to(device), self.W2.to(device), self.b2.to(device)
        return self

    def _weights_on(self, device):
        if self.W1.device != device:
            self.to(device)
        return self.W1, self.W2, self.b2

    def target_features(self, vol_tgt: torch.Tensor) -> torch.Tensor:
        """feature_aligner.forward_3d2d(img_feat_tgt) (modules/model.py:191)."""
        W1, W2, b2 = self._weights_on(vol_tgt.device)
        return ops.forward_3d2d(vol_tgt.float(), W1, W2, b2)

    def _workspace(self, B, N, k, device):
        need = ops.workspace_bytes(B, N, max(k, 1))
        if self._ws is None or self._ws.device != device or self._ws.numel() < need:
            self._ws = torch.empty(max(need, 16), device=device, dtype=torch.uint8)
        return self._ws

    @torch.no_grad()
    def score(self, vol_src, vol_tgt, R, k: int = 1, return_scores: bool = True, tgt_feat=None,
              idx_offset: int = 0, gather: bool = True) -> VerifyResult:
        """vol_src/vol_tgt [B,16,8,8,8]; R [N,3,3] shared or [B,N,3,3] per pair."""
        dev = vol_src.device
        W1, W2, b2 = self._weights_on(dev)
        per_pair = R.dim() == 4
        N = R.shape[1] if per_pair else R.shape[0]
        B = vol_src.shape[0]
        if N == 0:
            raise ValueError("empty hypothesis set")
        k = min(k, N)
        vs = vol_src if vol_src.dtype == torch.bfloat16 else vol_src.float()
        ws = self._workspace(B, N, k, dev)
        if tgt_feat is None:   # one C call for the whole step (2 launches when k == 1)
            scores, val, idx, R_best = ops.verify(vs, vol_tgt.float(), R, W1, W2, b2, k=k, idx_offset=idx_offset,
                                                  math=self.math, return_scores=return_scores, gather=gather,
                                                  workspace=ws)
            return VerifyResult(scores, val, idx, R_best)
        scores, val, idx = ops.score(vs, tgt_feat, R, W1, W2, b2, k=k, idx_offset=idx_offset, math=self.math,
                                     return_scores=return_scores, workspace=ws)
        R_best = ops.gather_rotations(R, idx, idx_offset) if gather else None
        return VerifyResult(scores, val, idx, R_best)

    @torch.no_grad()
    def predict(self, vol_src, vol_tgt, R):
        """argmax hypothesis per pair: (R_best [B,3,3], score [B], index [B])."""
        r = self.score(vol_src, vol_tgt, R, k=1, return_scores=False)
        return r.R_best[:, 0], r.topk_val[:, 0], r.topk_idx[:, 0]

    @torch.no_grad()
    def refine(self, vol_src, vol_tgt, R, k: int = 32, m: int = 64, max_angle_deg: float = 5.0, seed: int = 0):
        """Two-pass selection (BASELINE config 4; extension): score the set, keep the top-k, score m local
        perturbations of each, return the best - one C call (`ahv_refine`), no torch ops in between.
        Returns (R_best [B,3,3], score [B], first pass as a VerifyResult, candidate set [B,k*m,3,3])."""
        dev = vol_src.device
        W1, W2, b2 = self._weights_on(dev)
        per_pair = R.dim() == 4
        k = min(k, R.shape[1] if per_pair else R.shape[0])
        vs = vol_src if vol_src.dtype == torch.bfloat16 else vol_src.float()
        fv, fi, fR, cand, bv, bi, bR, self._ws_refine = ops.refine(vs, vol_tgt.float(), R, W1, W2, b2, k=k, m=m,
                                                                  max_angle_deg=max_angle_deg, seed=seed, math=self.math,
                                                                  workspace=getattr(self, "_ws_refine", None))
        return bR, bv, VerifyResult(None, fv, fi, fR), cand


class GraphedVerifier:
    """CUDA-graph replay of one verification step for fixed (B, N, k): target
    features + fused scoring + selection + winner gather are captured once and
    replayed with a single launch, which is what bounds the B=1 latency
    (the reference's per-pair loop is batch 1, test_co3d.py:133).  Inputs are
    copied into static buffers; outputs are static tensors valid until the next
    replay."""

    def __init__(self, verifier: HypothesisVerifier, B: int, N: int, k: int = 1, per_pair_R: bool = False,
                 device="cuda", vol_dtype=torch.float32, peer=None, idx_offset: int = 0):
        """`peer` (a `dist.PeerExchange`): N is THIS rank's slice of a hypothesis set sharded over the peer
        group (global index = local + idx_offset) and the captured step is `ahv_verify_sharded` - the scoring
        kernels exchange their winners over NVLink themselves, so the multi-GPU step contains no NCCL call and
        is graph-capturable as it stands.  Every rank must replay in lockstep."""
        self.v = verifier.to(torch.device(device))
        dev = torch.device(device)
        self.vol_src = torch.zeros(B, 16, 8, 8, 8, device=dev, dtype=vol_dtype)
        self.vol_tgt = torch.zeros(B, 16, 8, 8, 8, device=dev)
        self.R = torch.eye(3, device=dev).repeat(*((B, N) if per_pair_R else (N,)), 1, 1).contiguous()
        self.k = min(k, N)
        self.peer = peer

        def run():
            if peer is None:
                return self.v.score(self.vol_src, self.vol_tgt, self.R, k=self.k, return_scores=False, idx_offset=idx_offset)
            W1, W2, b2 = self.v._weights_on(dev)
            val, idx, Rb = ops.verify_sharded(self.vol_src, self.vol_tgt, self.R, W1, W2, b2, idx_offset, peer, k=self.k,
                                              math=self.v.math, workspace=self.v._workspace(B, N, self.k, dev))
            return VerifyResult(None, val, idx, Rb)

        stream = torch.cuda.Stream(device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream):
            for _ in range(2):                      # warm-up outside capture (module load, attributes)
                self.out = run()
        torch.cuda.current_stream(dev).wait_stream(stream)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = run()

    @torch.no_grad()
    def __call__(self, vol_src=None, vol_tgt=None, R=None) -> VerifyResult:
        if vol_src is not None:
            self.vol_src.copy_(vol_src, non_blocking=True)
        if vol_tgt is not None:
            self.vol_tgt.copy_(vol_tgt, non_blocking=True)
        if R is not None:
            self.R.copy_(R, non_blocking=True)
        self.graph.replay()
        return self.out

    def check(self) -> int:
        """Sharded replays only: synchronises and raises if an exchange timed out waiting for a peer (that step's
        results are NaN / index -1); returns the number of exchanges this rank has completed."""
        return self.peer.check() if self.peer is not None else 0


class GraphedRefiner:
    """CUDA-graph replay of the two-pass selection (`ahv_refine`: dense set -> top-k -> m local candidates each ->
    arg-max; BASELINE config 4) for fixed (B, N, k, m): the whole chain - target features, both scoring passes, top-k,
    candidate generation, selection - is one C call with no host step or torch op inside, so it captures as it stands.
    Outputs are static tensors valid until the next replay."""

    def __init__(self, verifier: HypothesisVerifier, B: int, N: int, k: int = 32, m: int = 64, max_angle_deg: float = 5.0,
                 seed: int = 0, device="cuda", vol_dtype=torch.float32):
        dev = torch.device(device)
        self.v = verifier.to(dev)
        self.vol_src = torch.zeros(B, 16, 8, 8, 8, device=dev, dtype=vol_dtype)
        self.vol_tgt = torch.zeros(B, 16, 8, 8, 8, device=dev)
        self.R = torch.eye(3, device=dev).repeat(N, 1, 1).contiguous()
        k = min(k, N)
        W1, W2, b2 = self.v._weights_on(dev)

        def run(out=None, ws=None):
            return ops.refine(self.vol_src, self.vol_tgt, self.R, W1, W2, b2, k=k, m=m, max_angle_deg=max_angle_deg, seed=seed,
                              math=self.v.math, workspace=ws, out=out)

        stream = torch.cuda.Stream(device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream):
            res = run()
            res = run(res[:7], res[7])             # warm-up on the buffers the graph will use
        torch.cuda.current_stream(dev).wait_stream(stream)
        torch.cuda.synchronize(dev)
        self._bufs, self._ws = res[:7], res[7]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            run(self._bufs, self._ws)
        fv, fi, fR, self.candidates, self.best_val, self.best_idx, self.R_best = self._bufs
        self.first = VerifyResult(None, fv, fi, fR)

    @torch.no_grad()
    def __call__(self, vol_src=None, vol_tgt=None, R=None):
        """Returns (R_best [B,3,3], score [B]); `.first` holds the first pass's top-k, `.candidates` the refinement set."""
        if vol_src is not None:
            self.vol_src.copy_(vol_src, non_blocking=True)
        if vol_tgt is not None:
            self.vol_tgt.copy_(vol_tgt, non_blocking=True)
        if R is not None:
            self.R.copy_(R, non_blocking=True)
        self.graph.replay()
        return self.R_best, self.best_val


class GraphedTail:
    """CUDA-graph replay of everything downstream of the 2D backbone for fixed (B, N, k): `forward_2d3d`
    (1x1 conv + 2D ResNet block, bidirectional transformer, 3D ResNet block - `ahv_resblock3d` - for both views;
    modules/modules.py:86-110) followed by the fused verification step, captured once and replayed with a single
    launch.  This is the per-pair latency path of the reference's evaluation loops (test_co3d.py:133-146) minus
    the backbone.  Inputs are copied into static buffers; outputs are static tensors valid until the next replay."""

    def __init__(self, feature_aligner, B: int, N: int, k: int = 1, device="cuda", math: int | None = None,
                 in_channels: int = 768):
        dev = torch.device(device)
        self.fa = feature_aligner
        self.v = HypothesisVerifier.from_feature_aligner(feature_aligner, math).to(dev)
        self.feat_src = torch.zeros(B, in_channels, 8, 8, device=dev)
        self.feat_tgt = torch.zeros(B, in_channels, 8, 8, device=dev)
        self.R = torch.eye(3, device=dev).repeat(N, 1, 1).contiguous()
        self.k = min(k, N)

        def run():
            with torch.no_grad():
                vs, vt = self.fa.forward_2d3d(self.feat_src, self.feat_tgt, random_mask=False, mask_ratio=0.0)
                return self.v.score(vs, vt, self.R, k=self.k, return_scores=False), vs, vt

        stream = torch.cuda.Stream(device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream):
            for _ in range(3):                      # warm-up outside capture (cuDNN/cuBLAS plans, module load)
                self.out, self.vol_src, self.vol_tgt = run()
        torch.cuda.current_stream(dev).wait_stream(stream)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out, self.vol_src, self.vol_tgt = run()

    @torch.no_grad()
    def __call__(self, feat_src=None, feat_tgt=None, R=None) -> VerifyResult:
        if feat_src is not None:
            self.feat_src.copy_(feat_src, non_blocking=True)
        if feat_tgt is not None:
            self.feat_tgt.copy_(feat_tgt, non_blocking=True)
        if R is not None:
            self.R.copy_(R, non_blocking=True)
        self.graph.replay()
        return self.out
