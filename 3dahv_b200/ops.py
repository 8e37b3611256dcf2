"""Tensor-level wrappers over the C ABI: device pointers, current stream, checks.

PyTorch is used only for device memory and streams; all arithmetic of the hot
path happens inside lib3dahv_b200.so.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import MATH_FP32, MATH_TC, MATH_TC_F16GATHER, VOL_BF16, VOL_F32

_BASE_CPU = None
_BASE_DEV = {}


def base_coords(device=None) -> torch.Tensor:
    """The 8 base coordinates F.affine_grid(align_corners=False) uses for size 8
    (ATen: linspace(-1,1,8)*7/8 — not exactly (2i+1)/8-1), taken from ATen itself
    so the kernel sees the reference's values (utils.py:126).  Cached per device
    (no host->device copy on the hot path, CUDA-graph safe)."""
    global _BASE_CPU
    if _BASE_CPU is None:
        g = F.affine_grid(torch.eye(3, 4)[None], (1, 1, 8, 8, 8), align_corners=False)
        _BASE_CPU = g[0, 0, 0, :, 0].contiguous().clone()
    if device is None:
        return _BASE_CPU
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if device not in _BASE_DEV:
        _BASE_DEV[device] = _BASE_CPU.to(device)
    return _BASE_DEV[device]


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _dev(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the 3DAHV hot path has no CPU fallback")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    t = t.contiguous()
    if t.data_ptr() % 16:   # the C ABI wants 16-byte aligned buffers; a slice such as R[lo:hi] need not be
        t = t.clone()
    return t


def rotations_from_normals(normals: torch.Tensor) -> torch.Tensor:
    """[n,4] normals (GPU) -> [n,3,3]; pytorch3d random_rotations arithmetic."""
    o = _dev(normals, "normals")
    n = o.shape[0]
    R = torch.empty(n, 3, 3, device=o.device, dtype=torch.float32)
    with torch.cuda.device(o.device):
        _lib.check(_lib.lib().ahv_so3_from_normals(o.data_ptr(), R.data_ptr(), n, _stream(o)), "ahv_so3_from_normals")
    return R


def sample_rotations(n: int, seed: int, first_index: int = 0, device="cuda") -> torch.Tensor:
    """Native Philox sampler: hypothesis i depends only on (seed, first_index+i)."""
    device = torch.device(device)
    R = torch.empty(n, 3, 3, device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().ahv_so3_sample(seed & (2**64 - 1), first_index, R.data_ptr(), n, _stream(R)), "ahv_so3_sample")
    return R


def grid_rotations(n_total: int, first_index: int = 0, count: int | None = None, device="cuda") -> torch.Tensor:
    """Deterministic super-Fibonacci SO(3) grid, points [first, first+count) of n_total."""
    device = torch.device(device)
    count = n_total - first_index if count is None else count
    R = torch.empty(count, 3, 3, device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().ahv_so3_grid(n_total, first_index, R.data_ptr(), count, _stream(R)), "ahv_so3_grid")
    return R


def perturb_rotations(R_center: torch.Tensor, m: int, max_angle_deg: float, seed: int = 0) -> torch.Tensor:
    """ahv_so3_perturb: [...,3,3] centres -> [...,m,3,3] rotations within `max_angle_deg` of them (index 0 = the
    centre itself)."""
    c = _dev(R_center, "R_center")
    lead = c.shape[:-2]
    n = c.numel() // 9
    out = torch.empty(*lead, m, 3, 3, device=c.device, dtype=torch.float32)
    with torch.cuda.device(c.device):
        _lib.check(_lib.lib().ahv_so3_perturb(c.data_ptr(), n, m, float(max_angle_deg), seed & (2**64 - 1), out.data_ptr(),
                                              _stream(c)), "ahv_so3_perturb")
    return out


def rotate_volume(volume: torch.Tensor, R: torch.Tensor) -> torch.Tensor:
    """utils.rotate_volume (utils.py:113-131) on the GPU.  `volume` is
    [16,8,8,8] (one volume under n rotations) or [n,16,8,8,8] (one per rotation)."""
    R = _dev(R, "R")
    n = R.shape[0]
    per_rot = volume.dim() == 5
    if per_rot and volume.shape[0] != n:
        raise ValueError("volume batch must match the number of rotations")
    if tuple(volume.shape[-4:]) != (16, 8, 8, 8):
        raise ValueError("volume must be [...,16,8,8,8]")
    v = _dev(volume, "volume")
    out = torch.empty(n, 16, 8, 8, 8, device=v.device, dtype=torch.float32)
    base = base_coords(v.device)
    with torch.cuda.device(v.device):
        _lib.check(_lib.lib().ahv_rotate_volume(v.data_ptr(), int(per_rot), R.data_ptr(), base.data_ptr(), out.data_ptr(), n, _stream(v)), "ahv_rotate_volume")
    return out


def rotate_volume_backward(grad_out: torch.Tensor, R: torch.Tensor, per_rotation: bool) -> torch.Tensor:
    """Adjoint of rotate_volume w.r.t. the volume: [n,16,8,8,8] -> [16,8,8,8] (shared volume, summed)
    or [n,16,8,8,8] (one volume per rotation)."""
    g, R = _dev(grad_out, "grad_out"), _dev(R, "R")
    n = R.shape[0]
    if tuple(g.shape) != (n, 16, 8, 8, 8):
        raise ValueError("grad_out must be [n,16,8,8,8]")
    out = (torch.empty(n, 16, 8, 8, 8, device=g.device, dtype=torch.float32) if per_rotation
           else torch.zeros(16, 8, 8, 8, device=g.device, dtype=torch.float32))
    base = base_coords(g.device)
    with torch.cuda.device(g.device):
        _lib.check(_lib.lib().ahv_rotate_volume_backward(g.data_ptr(), int(per_rotation), R.data_ptr(), base.data_ptr(),
                                                         out.data_ptr(), n, _stream(g)), "ahv_rotate_volume_backward")
    return out


def score_train(vol_src: torch.Tensor, tgt_feat: torch.Tensor, R: torch.Tensor, W1: torch.Tensor, W2: torch.Tensor,
                b2: torch.Tensor, workspace: torch.Tensor | None = None):
    """ahv_score_train: the training forward - scores [B,N] from the tensor-core kernel, plus what the backward pass
    wants back: conv1's ReLU'd output of every item (fp16 [B*N,64,32], 4 KB per item) and 1/scale per pair [B]."""
    vs, tf, R = _dev(vol_src, "vol_src"), _dev(tgt_feat, "tgt_feat"), _dev(R, "R")
    W1, W2, b2 = _dev(W1.reshape(32, 384), "W1"), _dev(W2.reshape(32, 32), "W2"), _dev(b2, "b2")
    B = vs.shape[0]
    per_pair = R.dim() == 4
    N = R.shape[1] if per_pair else R.shape[0]
    if tuple(vs.shape) != (B, 16, 8, 8, 8) or tuple(tf.shape) != (B, 32, 64) or (per_pair and R.shape[0] != B):
        raise ValueError("vol_src [B,16,8,8,8], tgt_feat [B,32,64], R [N,3,3] or [B,N,3,3] expected")
    dev = vs.device
    scores = torch.empty(B, N, device=dev, dtype=torch.float32)
    h1 = torch.empty(B * N, 64, 32, device=dev, dtype=torch.float16)
    pair_inv = torch.empty(B, device=dev, dtype=torch.float32)
    need = workspace_bytes(B, N, 1)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(max(need, 16), device=dev, dtype=torch.uint8)
    base = base_coords(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ahv_score_train(vs.data_ptr(), tf.data_ptr(), R.data_ptr(), int(per_pair), W1.data_ptr(), W2.data_ptr(),
                                              b2.data_ptr(), base.data_ptr(), scores.data_ptr(), h1.data_ptr(), pair_inv.data_ptr(),
                                              B, N, workspace.data_ptr(), workspace.numel() * workspace.element_size(),
                                              _stream(vs)), "ahv_score_train")
    return scores, h1, pair_inv


def score_backward(vol_src: torch.Tensor, tgt_feat: torch.Tensor, R: torch.Tensor, W1: torch.Tensor, W2: torch.Tensor,
                   b2: torch.Tensor, grad_scores: torch.Tensor, h1_saved: torch.Tensor | None = None,
                   pair_inv: torch.Tensor | None = None, math: int = MATH_TC):
    """Fused gradient of the verification scores (modules/model.py:53-56 under autograd): returns
    (grad_vol_src [B,16,8,8,8], grad_tgt_feat [B,32,64], grad_W1 [32,384], grad_W2 [32,32], grad_b2 [32]).
    With `h1_saved` / `pair_inv` from `score_train` the kernel reads conv1's output instead of recomputing it, and
    `math` picks how it contracts dA = dH1 W1 and dW1 = dH1^T A: MATH_TC on tcgen05, MATH_FP32 in FFMA.  Without them
    it is the exact fp32 kernel whatever `math` says."""
    vs, tg, R, gs = _dev(vol_src, "vol_src"), _dev(tgt_feat, "tgt_feat"), _dev(R, "R"), _dev(grad_scores, "grad_scores")
    W1, W2, b2 = _dev(W1.reshape(32, 384), "W1"), _dev(W2.reshape(32, 32), "W2"), _dev(b2, "b2")
    B = vs.shape[0]
    per_pair = R.dim() == 4
    N = R.shape[1] if per_pair else R.shape[0]
    if tuple(vs.shape) != (B, 16, 8, 8, 8) or tuple(tg.shape) != (B, 32, 64) or tuple(gs.shape) != (B, N):
        raise ValueError("vol_src [B,16,8,8,8], tgt_feat [B,32,64], grad_scores [B,N] expected")
    if per_pair and R.shape[0] != B:
        raise ValueError("per-pair rotations must be [B,N,3,3]")
    dev = vs.device
    z = lambda *shape: torch.zeros(*shape, device=dev, dtype=torch.float32)
    g_vol, g_tgt, g_w1, g_w2, g_b2 = z(B, 16, 8, 8, 8), z(B, 32, 64), z(32, 384), z(32, 32), z(32)
    base = base_coords(dev)
    with torch.cuda.device(dev):
        if h1_saved is not None:
            h1 = _dev(h1_saved, "h1_saved", torch.float16)
            pi = _dev(pair_inv, "pair_inv")
            if h1.numel() != B * N * 64 * 32 or pi.numel() != B:
                raise ValueError("h1_saved [B*N,64,32] fp16 and pair_inv [B] expected")
            _lib.check(_lib.lib().ahv_score_backward_saved(vs.data_ptr(), tg.data_ptr(), R.data_ptr(), int(per_pair), W1.data_ptr(),
                                                           W2.data_ptr(), b2.data_ptr(), base.data_ptr(), gs.data_ptr(),
                                                           h1.data_ptr(), pi.data_ptr(), g_vol.data_ptr(), g_tgt.data_ptr(),
                                                           g_w1.data_ptr(), g_w2.data_ptr(), g_b2.data_ptr(), B, N, int(math), _stream(vs)),
                       "ahv_score_backward_saved")
        else:
            _lib.check(_lib.lib().ahv_score_backward(vs.data_ptr(), tg.data_ptr(), R.data_ptr(), int(per_pair), W1.data_ptr(),
                                                     W2.data_ptr(), b2.data_ptr(), base.data_ptr(), gs.data_ptr(),
                                                     g_vol.data_ptr(), g_tgt.data_ptr(), g_w1.data_ptr(), g_w2.data_ptr(),
                                                     g_b2.data_ptr(), B, N, _stream(vs)), "ahv_score_backward")
    return g_vol, g_tgt, g_w1, g_w2, g_b2


def infonce(scores: torch.Tensor, sampled_R: torch.Tensor, gt_delta_R: torch.Tensor, acc_thr_deg: float,
            temperature: float = 0.1, want_grad: bool = True):
    """ahv_infonce: (loss [B], d loss / d scores [B,N] or None) in one launch (modules/model.py:43-63)."""
    s, R, g = _dev(scores, "scores"), _dev(sampled_R, "sampled_R"), _dev(gt_delta_R, "gt_delta_R")
    B, N = s.shape
    per_pair = R.dim() == 4
    if (R.shape[1] if per_pair else R.shape[0]) != N or tuple(g.shape) != (B, 3, 3) or (per_pair and R.shape[0] != B):
        raise ValueError("scores [B,N], sampled_R [N,3,3] or [B,N,3,3], gt_delta_R [B,3,3] expected")
    loss = torch.empty(B, device=s.device, dtype=torch.float32)
    grad = torch.empty(B, N, device=s.device, dtype=torch.float32) if want_grad else None
    with torch.cuda.device(s.device):
        _lib.check(_lib.lib().ahv_infonce(s.data_ptr(), R.data_ptr(), int(per_pair), g.data_ptr(), float(acc_thr_deg),
                                          float(temperature), loss.data_ptr(), grad.data_ptr() if want_grad else None, B, N,
                                          _stream(s)), "ahv_infonce")
    return loss, grad


def resblock3d(x: torch.Tensor, conv1_w: torch.Tensor, conv2_w: torch.Tensor, down_w: torch.Tensor) -> torch.Tensor:
    """ResNetBlock_3D(32->16) of the lifting stage (modules/modules.py:9-47, :100-101): [m,32,8,8,8] -> [m,16,8,8,8]
    in one cluster launch (inference only)."""
    v = _dev(x, "x")
    if tuple(v.shape[1:]) != (32, 8, 8, 8):
        raise ValueError("x must be [m,32,8,8,8]")
    w1, w2, wd = _dev(conv1_w, "conv1_w"), _dev(conv2_w, "conv2_w"), _dev(down_w, "down_w")
    if tuple(w1.shape) != (16, 32, 3, 3, 3) or tuple(w2.shape) != (16, 16, 3, 3, 3) or w1.numel() * 0 + wd.numel() != 16 * 32:
        raise ValueError("weights must be conv1 [16,32,3,3,3], conv2 [16,16,3,3,3], downsample [16,32,1,1,1]")
    m = v.shape[0]
    out = torch.empty(m, 16, 8, 8, 8, device=v.device, dtype=torch.float32)
    with torch.cuda.device(v.device):
        _lib.check(_lib.lib().ahv_resblock3d(v.data_ptr(), w1.data_ptr(), w2.data_ptr(), wd.data_ptr(), out.data_ptr(), m,
                                             _stream(v)), "ahv_resblock3d")
    return out


def forward_3d2d(vol: torch.Tensor, W1: torch.Tensor, W2: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    """Feature_Aligner.forward_3d2d (modules/modules.py:112-124): [m,16,8,8,8] -> [m,32,64]."""
    v = _dev(vol, "vol")
    if tuple(v.shape[1:]) != (16, 8, 8, 8):
        raise ValueError("vol must be [m,16,8,8,8]")
    W1, W2, b2 = _dev(W1.reshape(32, 384), "W1"), _dev(W2.reshape(32, 32), "W2"), _dev(b2, "b2")
    m = v.shape[0]
    feat = torch.empty(m, 32, 64, device=v.device, dtype=torch.float32)
    with torch.cuda.device(v.device):
        _lib.check(_lib.lib().ahv_forward_3d2d(v.data_ptr(), W1.data_ptr(), W2.data_ptr(), b2.data_ptr(), feat.data_ptr(), m, _stream(v)), "ahv_forward_3d2d")
    return feat


def workspace_bytes(B: int, N: int, k: int) -> int:
    return int(_lib.lib().ahv_workspace_bytes(B, N, k))


def score(vol_src, tgt_feat, R, W1, W2, b2, k: int = 1, idx_offset: int = 0, math: int = MATH_TC,
          return_scores: bool = True, workspace: torch.Tensor | None = None):
    """The fused hot path (modules/model.py:186-196).

    vol_src [B,16,8,8,8] fp32|bf16, tgt_feat [B,32,64], R [N,3,3] or [B,N,3,3].
    Returns (scores [B,N] or None, topk_val [B,k], topk_idx [B,k] int64)."""
    if vol_src.dtype == torch.bfloat16:
        vs, vdt = _dev(vol_src, "vol_src", torch.bfloat16), VOL_BF16
    else:
        vs, vdt = _dev(vol_src, "vol_src"), VOL_F32
    B = vs.shape[0]
    if tuple(vs.shape[1:]) != (16, 8, 8, 8):
        raise ValueError("vol_src must be [B,16,8,8,8]")
    tf = _dev(tgt_feat, "tgt_feat")
    if tuple(tf.shape) != (B, 32, 64):
        raise ValueError("tgt_feat must be [B,32,64]")
    R = _dev(R, "R")
    per_pair = R.dim() == 4
    if per_pair and R.shape[0] != B:
        raise ValueError("per-pair R must be [B,N,3,3]")
    if tuple(R.shape[-2:]) != (3, 3):
        raise ValueError("R must be [...,3,3]")
    N = R.shape[1] if per_pair else R.shape[0]
    if k < 0 or k > 32:
        raise ValueError("k must be in [0,32]")
    W1, W2, b2 = _dev(W1.reshape(32, 384), "W1"), _dev(W2.reshape(32, 32), "W2"), _dev(b2, "b2")
    dev = vs.device
    scores = torch.empty(B, N, device=dev, dtype=torch.float32) if return_scores else None
    kk = max(k, 1)
    val = torch.empty(B, kk, device=dev, dtype=torch.float32)
    idx = torch.empty(B, kk, device=dev, dtype=torch.int64)
    need = workspace_bytes(B, N, kk)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(max(need, 16), device=dev, dtype=torch.uint8)
    base = base_coords(dev)
    with torch.cuda.device(dev):
        st = _lib.lib().ahv_score(
            vs.data_ptr(), vdt, tf.data_ptr(), R.data_ptr(), int(per_pair), W1.data_ptr(), W2.data_ptr(),
            b2.data_ptr(), base.data_ptr(), scores.data_ptr() if scores is not None else None,
            val.data_ptr() if k > 0 else None, idx.data_ptr() if k > 0 else None, k, idx_offset, B, N,
            math, workspace.data_ptr(), workspace.numel() * workspace.element_size(), _stream(vs))
    _lib.check(st, "ahv_score")
    if k == 0:
        return scores, None, None
    return scores, val, idx


def verify(vol_src, vol_tgt, R, W1, W2, b2, k: int = 1, idx_offset: int = 0, math: int = MATH_TC,
           return_scores: bool = False, gather: bool = True, workspace: torch.Tensor | None = None):
    """ahv_verify: target features + fused scoring + selection + winner gather in one C call.
    Returns (scores|None, topk_val [B,k], topk_idx [B,k], R_best [B,k,3,3]|None)."""
    if vol_src.dtype == torch.bfloat16:
        vs, vdt = _dev(vol_src, "vol_src", torch.bfloat16), VOL_BF16
    else:
        vs, vdt = _dev(vol_src, "vol_src"), VOL_F32
    vt = _dev(vol_tgt, "vol_tgt")
    B = vs.shape[0]
    if tuple(vs.shape[1:]) != (16, 8, 8, 8) or tuple(vt.shape) != (B, 16, 8, 8, 8):
        raise ValueError("vol_src / vol_tgt must be [B,16,8,8,8]")
    R = _dev(R, "R")
    per_pair = R.dim() == 4
    if per_pair and R.shape[0] != B:
        raise ValueError("per-pair R must be [B,N,3,3]")
    if tuple(R.shape[-2:]) != (3, 3):
        raise ValueError("R must be [...,3,3]")
    N = R.shape[1] if per_pair else R.shape[0]
    if k < 1 or k > 32:
        raise ValueError("k must be in [1,32]")
    W1, W2, b2 = _dev(W1.reshape(32, 384), "W1"), _dev(W2.reshape(32, 32), "W2"), _dev(b2, "b2")
    dev = vs.device
    scores = torch.empty(B, N, device=dev, dtype=torch.float32) if return_scores else None
    val = torch.empty(B, k, device=dev, dtype=torch.float32)
    idx = torch.empty(B, k, device=dev, dtype=torch.int64)
    Rb = torch.empty(B, k, 3, 3, device=dev, dtype=torch.float32) if gather else None
    need = workspace_bytes(B, N, k)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(max(need, 16), device=dev, dtype=torch.uint8)
    base = base_coords(dev)
    with torch.cuda.device(dev):
        st = _lib.lib().ahv_verify(
            vs.data_ptr(), vdt, vt.data_ptr(), R.data_ptr(), int(per_pair), W1.data_ptr(), W2.data_ptr(), b2.data_ptr(),
            base.data_ptr(), scores.data_ptr() if scores is not None else None, val.data_ptr(), idx.data_ptr(),
            Rb.data_ptr() if Rb is not None else None, k, idx_offset, B, N, math, workspace.data_ptr(),
            workspace.numel() * workspace.element_size(), _stream(vs))
    _lib.check(st, "ahv_verify")
    return scores, val, idx, Rb


def refine(vol_src, vol_tgt, R, W1, W2, b2, k: int = 32, m: int = 64, max_angle_deg: float = 5.0, seed: int = 0,
           math: int = MATH_TC, workspace: torch.Tensor | None = None, out=None):
    """ahv_refine: top-k over the set, m local candidates around each winner, arg-max over them - one C call, no
    torch ops, CUDA-graph capturable (pass `out` = the tuple a previous call returned to reuse its buffers).
    Returns (first_val [B,k], first_idx [B,k], first_R [B,k,3,3], cand [B,k*m,3,3], best_val [B], best_idx [B],
    R_best [B,3,3], workspace)."""
    if vol_src.dtype == torch.bfloat16:
        vs, vdt = _dev(vol_src, "vol_src", torch.bfloat16), VOL_BF16
    else:
        vs, vdt = _dev(vol_src, "vol_src"), VOL_F32
    vt, R = _dev(vol_tgt, "vol_tgt"), _dev(R, "R")
    B = vs.shape[0]
    per_pair = R.dim() == 4
    N = R.shape[1] if per_pair else R.shape[0]
    if tuple(vs.shape[1:]) != (16, 8, 8, 8) or tuple(vt.shape) != (B, 16, 8, 8, 8):
        raise ValueError("vol_src / vol_tgt must be [B,16,8,8,8]")
    if not (1 <= k <= min(32, N)) or m < 1:
        raise ValueError("1 <= k <= min(32, N) and m >= 1 expected")
    W1, W2, b2 = _dev(W1.reshape(32, 384), "W1"), _dev(W2.reshape(32, 32), "W2"), _dev(b2, "b2")
    dev = vs.device
    if out is None:
        f = dict(device=dev, dtype=torch.float32)
        out = (torch.empty(B, k, **f), torch.empty(B, k, device=dev, dtype=torch.int64), torch.empty(B, k, 3, 3, **f),
               torch.empty(B, k * m, 3, 3, **f), torch.empty(B, **f), torch.empty(B, device=dev, dtype=torch.int64),
               torch.empty(B, 3, 3, **f))
    fv, fi, fR, cand, bv, bi, bR = out[:7]
    need = int(_lib.lib().ahv_refine_workspace_bytes(B, N, k, m))
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(max(need, 16), device=dev, dtype=torch.uint8)
    base = base_coords(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ahv_refine(
            vs.data_ptr(), vdt, vt.data_ptr(), R.data_ptr(), int(per_pair), W1.data_ptr(), W2.data_ptr(), b2.data_ptr(),
            base.data_ptr(), k, m, float(max_angle_deg), seed & (2**64 - 1), fv.data_ptr(), fi.data_ptr(), fR.data_ptr(),
            cand.data_ptr(), bv.data_ptr(), bi.data_ptr(), bR.data_ptr(), B, N, math, workspace.data_ptr(),
            workspace.numel() * workspace.element_size(), _stream(vs)), "ahv_refine")
    return fv, fi, fR, cand, bv, bi, bR, workspace


def _peer_array(peer):
    import ctypes

    return (ctypes.c_void_p * peer.world)(*[int(p) for p in peer.ptrs])


def verify_sharded(vol_src, vol_tgt, R, W1, W2, b2, idx_offset: int, peer, k: int = 1, math: int = MATH_TC,
                   workspace: torch.Tensor | None = None):
    """ahv_verify_sharded on this rank's slice `R` of the rotation set (global index = local + idx_offset); `peer`
    is a `dist.PeerExchange`.  The kernels exchange the per-shard winners through NVLink peer memory and merge
    them (k == 1 with tensor-core arithmetic: inside the scoring kernel; otherwise one exchange kernel after the
    shard's top-k).  Returns (topk_val [B,k], topk_idx [B,k] global, R_best [B,k,3,3]) over the whole set,
    identical on every rank.  An empty slice ([0,3,3]) is allowed."""
    vs = _dev(vol_src, "vol_src", vol_src.dtype if vol_src.dtype == torch.bfloat16 else torch.float32)
    vt, R = _dev(vol_tgt, "vol_tgt"), _dev(R, "R")
    W1, W2, b2 = _dev(W1.reshape(32, 384), "W1"), _dev(W2.reshape(32, 32), "W2"), _dev(b2, "b2")
    B = vs.shape[0]
    per_pair = R.dim() == 4
    N = R.shape[1] if per_pair else R.shape[0]
    if tuple(vs.shape[1:]) != (16, 8, 8, 8) or tuple(vt.shape) != (B, 16, 8, 8, 8):
        raise ValueError("vol_src / vol_tgt must be [B,16,8,8,8]")
    if B > peer.max_pairs or k > peer.max_k:
        raise ValueError(f"exchange buffer holds {peer.max_pairs} pairs x top-{peer.max_k}; got B={B}, k={k}")
    dev = vs.device
    val = torch.empty(B, k, device=dev, dtype=torch.float32)
    idx = torch.empty(B, k, device=dev, dtype=torch.int64)
    Rb = torch.empty(B, k, 3, 3, device=dev, dtype=torch.float32)
    need = workspace_bytes(B, N, k)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(max(need, 16), device=dev, dtype=torch.uint8)
    base = base_coords(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ahv_verify_sharded(
            vs.data_ptr(), VOL_BF16 if vs.dtype == torch.bfloat16 else VOL_F32, vt.data_ptr(), R.data_ptr() if N else None,
            int(per_pair), W1.data_ptr(), W2.data_ptr(), b2.data_ptr(), base.data_ptr(), val.data_ptr(), idx.data_ptr(),
            Rb.data_ptr(), k, idx_offset, B, N, math, workspace.data_ptr(), workspace.numel() * workspace.element_size(),
            peer.rank, peer.world, _peer_array(peer), peer.max_pairs, peer.max_k, _stream(vs)), "ahv_verify_sharded")
    return val, idx, Rb


def topk_exchange(val, idx, R, idx_offset: int, peer):
    """ahv_topk_exchange: this shard's [B,k] list (global indices, -1 = empty) -> the whole set's top-k on every
    rank, with the winning rotations.  Returns (topk_val [B,k], topk_idx [B,k], R_best [B,k,3,3])."""
    v, i, R = _dev(val, "val"), _dev(idx, "idx", torch.int64), _dev(R, "R")
    B, k = v.shape
    per_pair = R.dim() == 4
    N = R.shape[1] if per_pair else R.shape[0]
    dev = v.device
    out_v = torch.empty(B, k, device=dev, dtype=torch.float32)
    out_i = torch.empty(B, k, device=dev, dtype=torch.int64)
    Rb = torch.empty(B, k, 3, 3, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ahv_topk_exchange(v.data_ptr(), i.data_ptr(), R.data_ptr() if N else None, int(per_pair), idx_offset,
                                                N, B, k, out_v.data_ptr(), out_i.data_ptr(), Rb.data_ptr(), peer.rank, peer.world,
                                                _peer_array(peer), peer.max_pairs, peer.max_k, _stream(v)), "ahv_topk_exchange")
    return out_v, out_i, Rb


def topk(scores: torch.Tensor, k: int, idx_offset: int = 0):
    """torch.max / top-k with the reference's tie rule (lowest index wins)."""
    s = _dev(scores, "scores")
    B, N = s.shape
    val = torch.empty(B, k, device=s.device, dtype=torch.float32)
    idx = torch.empty(B, k, device=s.device, dtype=torch.int64)
    need = workspace_bytes(B, N, k)
    ws = torch.empty(max(need, 16), device=s.device, dtype=torch.uint8)
    with torch.cuda.device(s.device):
        st = _lib.lib().ahv_topk(s.data_ptr(), B, N, k, idx_offset, val.data_ptr(), idx.data_ptr(), ws.data_ptr(), ws.numel(), _stream(s))
    _lib.check(st, "ahv_topk")
    return val, idx


def topk_merge(vals: torch.Tensor, idx: torch.Tensor):
    """Merge [parts,B,k] shard lists into [B,k] (same ordering rule)."""
    v, i = _dev(vals, "vals"), _dev(idx, "idx", torch.int64)
    parts, B, k = v.shape
    out_v = torch.empty(B, k, device=v.device, dtype=torch.float32)
    out_i = torch.empty(B, k, device=v.device, dtype=torch.int64)
    with torch.cuda.device(v.device):
        st = _lib.lib().ahv_topk_merge(v.data_ptr(), i.data_ptr(), parts, B, k, out_v.data_ptr(), out_i.data_ptr(), _stream(v))
    _lib.check(st, "ahv_topk_merge")
    return out_v, out_i


def gather_rotations(R: torch.Tensor, idx: torch.Tensor, idx_offset: int = 0) -> torch.Tensor:
    """sampled_R[pred_index] (modules/model.py:196): [B,k,3,3]."""
    R, idx = _dev(R, "R"), _dev(idx, "idx", torch.int64)
    per_pair = R.dim() == 4
    B, k = idx.shape
    N = R.shape[1] if per_pair else R.shape[0]
    out = torch.empty(B, k, 3, 3, device=R.device, dtype=torch.float32)
    with torch.cuda.device(R.device):
        st = _lib.lib().ahv_gather_rotations(R.data_ptr(), int(per_pair), idx.data_ptr(), idx_offset, B, N, k, out.data_ptr(), _stream(R))
    _lib.check(st, "ahv_gather_rotations")
    return out


class HostSession:
    """Caller-owned device scratch of the host-buffer entry (`ahv_host_session_create`), reused across calls."""

    def __init__(self, device="cuda"):
        import ctypes

        self.device = torch.device(device)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ahv_host_session_create(ctypes.byref(h)), "ahv_host_session_create")
        self.handle = h

    def close(self):
        if getattr(self, "handle", None) is not None:
            with torch.cuda.device(self.device):
                _lib.lib().ahv_host_session_destroy(self.handle)
            self.handle = None

    __del__ = close


def predict_host(vol_src, vol_tgt, R, W1, W2, b2, k: int = 1, math: int = MATH_TC, return_scores: bool = False,
                 device="cuda", session: HostSession | None = None, idx_offset: int = 0, peer=None):
    """HOST tensors in, HOST results out (copies inside the call): `ahv_predict_host`, or with a `session`
    (reused device scratch), bf16 source volumes, an index offset or a `peer` exchange (`R` = this rank's slice
    of a sharded rotation set; results over the whole set) `ahv_predict_host_ex`."""
    f = torch.float32
    bf16 = vol_src.dtype == torch.bfloat16
    vs, vt = (vol_src if bf16 else vol_src.to(f)).contiguous(), vol_tgt.to(f).contiguous()
    Rc = R.to(f).contiguous()
    if vs.is_cuda or vt.is_cuda or Rc.is_cuda:
        raise RuntimeError("predict_host takes host tensors")
    W1c, W2c, b2c = W1.reshape(32, 384).to(f).contiguous().cpu(), W2.reshape(32, 32).to(f).contiguous().cpu(), b2.to(f).contiguous().cpu()
    per_pair = Rc.dim() == 4
    B = vs.shape[0]
    N = Rc.shape[1] if per_pair else Rc.shape[0]
    scores = torch.empty(B, N, dtype=f) if return_scores else None
    val = torch.empty(B, k, dtype=f)
    idx = torch.empty(B, k, dtype=torch.int64)
    Rb = torch.empty(B, k, 3, 3, dtype=f)
    base = base_coords()
    device = torch.device(device)
    extended = session is not None or bf16 or idx_offset != 0 or peer is not None
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        if not extended:
            st = _lib.lib().ahv_predict_host(
                vs.data_ptr(), vt.data_ptr(), Rc.data_ptr(), int(per_pair), W1c.data_ptr(), W2c.data_ptr(), b2c.data_ptr(),
                base.data_ptr(), scores.data_ptr() if scores is not None else None, val.data_ptr(), idx.data_ptr(),
                Rb.data_ptr(), k, B, N, math, stream)
        else:
            own = session is None
            if own:
                session = HostSession(device)
            try:
                st = _lib.lib().ahv_predict_host_ex(
                    session.handle, vs.data_ptr(), VOL_BF16 if bf16 else VOL_F32, vt.data_ptr(), Rc.data_ptr(), int(per_pair),
                    W1c.data_ptr(), W2c.data_ptr(), b2c.data_ptr(), base.data_ptr(),
                    scores.data_ptr() if scores is not None else None, val.data_ptr(), idx.data_ptr(), Rb.data_ptr(), k,
                    idx_offset, B, N, math, peer.rank if peer else 0, peer.world if peer else 1,
                    _peer_array(peer) if peer else None, peer.max_pairs if peer else 0, peer.max_k if peer else 0, stream)
            finally:
                if own:
                    session.close()
    _lib.check(st, "ahv_predict_host")
    return scores, val, idx, Rb
