"""SO(3) hypothesis sets.

`random_rotations` mirrors `pytorch3d.transforms.random_rotations` as the
reference calls it (modules/model.py:102,131,184; model_co3d.py:86;
test_co3d.py:106): the Gaussian draws come from torch's CPU generator — the
reference passes no device, so a GPU Philox stream could never reproduce its
hypothesis set — and only the cheap quaternion -> matrix map runs on the GPU
(`ahv_so3_from_normals`, 16 B uploaded per hypothesis instead of 36 B).
"""
from __future__ import annotations

import torch

from . import ops


def draw_normals(n: int, generator: torch.Generator | None = None) -> torch.Tensor:
    """`torch.randn((n,4))` on the CPU default (or given) generator, exactly the
    draw pytorch3d's random_quaternions makes."""
    return torch.randn((n, 4), dtype=torch.float32, generator=generator)


def random_rotations(n: int, dtype=None, device=None, generator: torch.Generator | None = None) -> torch.Tensor:
    """Haar-uniform rotations [n,3,3], hypothesis set identical to the reference's
    for the same CPU RNG state.  Result lives on `device` (default: current CUDA
    device, as the reference moves it there right away, modules/model.py:184)."""
    if dtype not in (None, torch.float32):
        raise TypeError("the 3DAHV path samples rotations in float32")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("random_rotations computes on the GPU; pass a CUDA device (no CPU fallback)")
    normals = draw_normals(n, generator)
    return ops.rotations_from_normals(normals.to(dev, non_blocking=True))


def sample_rotations(n: int, seed: int = 0, first_index: int = 0, device="cuda") -> torch.Tensor:
    """Native counter-based sampler (Philox-4x32-10): any shard [first, first+n)
    of the set can be generated on any GPU with no host step."""
    return ops.sample_rotations(n, seed, first_index, device)


def grid_rotations(n_total: int, first_index: int = 0, count: int | None = None, device="cuda") -> torch.Tensor:
    """Deterministic, near-uniform SO(3) grid of any size (super-Fibonacci spiral): the
    "dense SO(3) grid" of BASELINE config 4.  Shardable like the native sampler."""
    return ops.grid_rotations(n_total, first_index, count, device)


def perturb_rotations(R_center: torch.Tensor, m: int, max_angle_deg: float, seed: int = 0) -> torch.Tensor:
    """Local refinement set (BASELINE config 4; extension, the reference has no
    refinement): for each of the [...,3,3] centres, m rotations within
    `max_angle_deg` of it (index 0 is the centre itself).  Built from the native
    sampler: a Haar rotation's axis with the angle rescaled into the cone."""
    dev = R_center.device
    lead = R_center.shape[:-2]
    c = R_center.reshape(-1, 3, 3)
    n = c.shape[0]
    S = ops.sample_rotations(n * m, seed, 0, dev).reshape(n, m, 3, 3)
    # axis-angle of S, rescale angle to [0, max_angle]
    tr = S.diagonal(dim1=-2, dim2=-1).sum(-1)
    ang = torch.arccos(((tr - 1) / 2).clamp(-1, 1))
    axis = torch.stack([S[..., 2, 1] - S[..., 1, 2], S[..., 0, 2] - S[..., 2, 0], S[..., 1, 0] - S[..., 0, 1]], -1)
    axis = axis / axis.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    new_ang = ang / torch.pi * (max_angle_deg * torch.pi / 180.0)
    K = torch.zeros(n, m, 3, 3, device=dev)
    K[..., 0, 1], K[..., 0, 2] = -axis[..., 2], axis[..., 1]
    K[..., 1, 0], K[..., 1, 2] = axis[..., 2], -axis[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -axis[..., 1], axis[..., 0]
    s, co = torch.sin(new_ang)[..., None, None], torch.cos(new_ang)[..., None, None]
    dR = torch.eye(3, device=dev) + s * K + (1 - co) * (K @ K)
    dR[:, 0] = torch.eye(3, device=dev)
    out = dR @ c[:, None]
    return out.reshape(*lead, m, 3, 3)
