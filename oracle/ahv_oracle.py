"""CPU oracle for the 3DAHV hypothesis-and-verification hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (`3dahv_b200/`) may
import this module; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` use it, and only as the
checker or the CPU baseline, never as the thing shipped.

Parity status: PINNED.  Every function here is checked by
`tests/test_oracle_golden.py` against `tests/golden/*.npz`, which were written
by `oracle/make_golden.py` running the reference's OWN code
(`/root/reference/utils.py:113-131 rotate_volume`,
`/root/reference/modules/modules.py:112-124 Feature_Aligner.forward_3d2d`,
and the scoring idiom `/root/reference/modules/model.py:186-196`) in the build
container.  The one exception is `random_rotations`: pytorch3d is a third-party
dependency that is absent from `/root/reference` and from this image (unpinned,
`install.sh:9`), so its published algorithm is restated here and the parity of
that single function is "unpinned" (structural checks only: det=+1,
orthogonality, real part >= 0).

Two restatements are provided:

* `*_torch`  — the reference idiom re-expressed with the same ATen calls the
  reference makes (`F.affine_grid`, `F.grid_sample`, 1x1 convs, `F.normalize`),
  chunked over hypotheses so it fits in memory.  On the same torch build it is
  bit-identical to the reference; it is also the "port" CPU baseline that
  `bench.py` times on all host threads.
* `*_np`     — explicit scalar arithmetic in numpy (coordinates from the 3x3
  rotation, 8 individually bounds-checked taps, tri-plane contraction, ReLU,
  second contraction, L2 normalise, dot, mean).  This is the arithmetic the CUDA
  kernel implements, written without any library resampler.
"""
from __future__ import annotations

import math
import numpy as np

C, D, H, W = 16, 8, 8, 8          # modules/modules.py:64,97-100
VOX = D * H * W                   # 512 voxel samples per hypothesis
KTRI = 3 * 8 * C                  # 384, modules/modules.py:67
OCH = 32                          # out_channel, modules/model.py:34
NPOS = 64                         # 8x8 positions after the tri-plane fold


# --------------------------------------------------------------------------
# a1: pytorch3d.transforms.random_rotations (third-party, restated)
# call sites: modules/model.py:102,131,184  modules/model_co3d.py:86
#             test_co3d.py:106  test_linemod.py:43
# --------------------------------------------------------------------------
def rotations_from_normals_np(o: np.ndarray) -> np.ndarray:
    """[n,4] fp32 Gaussian draws -> [n,3,3] fp32 rotation matrices.

    Follows pytorch3d 0.7.x `random_quaternions` + `quaternion_to_matrix`:
    every product / sum is an individually rounded fp32 op, sums over the 4
    components are sequential ((a+b)+c)+d  ([probe]).
    """
    o = np.ascontiguousarray(o, dtype=np.float32)
    f = np.float32
    sq = o * o
    s = ((sq[:, 0] + sq[:, 1]) + sq[:, 2]) + sq[:, 3]
    den = np.copysign(np.sqrt(s).astype(f), o[:, 0]).astype(f)
    q = (o / den[:, None]).astype(f)
    r, i, j, k = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    qq = q * q
    two_s = (f(2.0) / (((qq[:, 0] + qq[:, 1]) + qq[:, 2]) + qq[:, 3])).astype(f)
    one = f(1.0)
    R = np.stack(
        (
            one - two_s * (j * j + k * k),
            two_s * (i * j - k * r),
            two_s * (i * k + j * r),
            two_s * (i * j + k * r),
            one - two_s * (i * i + k * k),
            two_s * (j * k - i * r),
            two_s * (i * k - j * r),
            two_s * (j * k + i * r),
            one - two_s * (i * i + j * j),
        ),
        axis=-1,
    ).astype(f)
    return R.reshape(-1, 3, 3)


def rotations_from_normals_torch(o):
    """Same as above with torch CPU ops in the order pytorch3d issues them."""
    import torch

    s = (o * o).sum(1)
    q = o / torch.copysign(torch.sqrt(s), o[:, 0])[:, None]
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    R = torch.stack(
        (
            1 - two_s * (j * j + k * k),
            two_s * (i * j - k * r),
            two_s * (i * k + j * r),
            two_s * (i * j + k * r),
            1 - two_s * (i * i + k * k),
            two_s * (j * k - i * r),
            two_s * (i * k - j * r),
            two_s * (j * k + i * r),
            1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return R.reshape(-1, 3, 3)


def random_rotations_torch(n: int, generator=None):
    """`random_rotations(n)`: normals from torch's CPU generator (the reference
    passes no device, modules/model.py:184), then quaternion -> matrix."""
    import torch

    o = torch.randn((n, 4), dtype=torch.float32, generator=generator)
    return rotations_from_normals_torch(o)


# --------------------------------------------------------------------------
# upstream of the path: ResNetBlock_3D(32 -> 16, BN=False, stride 1)
# modules/modules.py:9-47 as applied at :100-101
# --------------------------------------------------------------------------
def _conv3d_np(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """Conv3d, kernel 3, padding 1, stride 1, no bias: x [m,ci,8,8,8], w [co,ci,3,3,3] (float64 accumulation)."""
    m, ci, D, H, W_ = x.shape
    xp = np.zeros((m, ci, D + 2, H + 2, W_ + 2), dtype=np.float64)
    xp[:, :, 1:-1, 1:-1, 1:-1] = x
    out = np.zeros((m, w.shape[0], D, H, W_), dtype=np.float64)
    for kd in range(3):
        for kh in range(3):
            for kw in range(3):
                out += np.einsum("oc,mcdhw->modhw", w[:, :, kd, kh, kw].astype(np.float64),
                                 xp[:, :, kd:kd + D, kh:kh + H, kw:kw + W_])
    return out


def resblock3d_np(x, conv1_w, conv2_w, down_w) -> np.ndarray:
    """out = conv2(relu(conv1(x))) + downsample(x)   (modules/modules.py:32-47, bn1/bn2 empty, stride 1)."""
    h1 = np.maximum(_conv3d_np(np.asarray(x), np.asarray(conv1_w)).astype(np.float32), 0.0)
    out = _conv3d_np(h1, np.asarray(conv2_w))
    out += np.einsum("oc,mcdhw->modhw", np.asarray(down_w).reshape(down_w.shape[0], -1).astype(np.float64), np.asarray(x, dtype=np.float64))
    return out.astype(np.float32)


# --------------------------------------------------------------------------
# extensions beyond the reference: the native Philox sampler and
# the local refinement set, restated so the CUDA generators have an independent check
# (integer part exact; the transcendental functions agree to rounding, so the
# comparison is to ~1e-6, not bit-exact)
# --------------------------------------------------------------------------
def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox-4x32-10 (Salmon et al. 2011) on uint32 numpy arrays."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0 = np.uint32(k0) + np.zeros_like(c0)
    k1 = np.uint32(k1) + np.zeros_like(c0)
    m32 = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & m32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & m32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = k0 + W0
            k1 = k1 + W1
    return c0, c1, c2, c3


def _philox_quaternions(counter: np.ndarray, seed: int, s2: int, s3: int):
    """Four standard normals per counter (two Box-Muller pairs), as csrc/ahv_so3.cu draws them."""
    f = np.float32
    counter = np.asarray(counter, dtype=np.uint64)
    u = _philox4x32_10((counter & np.uint64(0xFFFFFFFF)).astype(np.uint32), (counter >> np.uint64(32)).astype(np.uint32),
                       np.full(counter.shape, s2, np.uint32), np.full(counter.shape, s3, np.uint32),
                       seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    inv32 = f(2.3283064365386963e-10)
    u0 = ((u[0] >> np.uint32(8)).astype(f) + f(0.5)) * f(1.0 / 16777216.0)
    u1 = u[1].astype(f) * inv32
    u2 = ((u[2] >> np.uint32(8)).astype(f) + f(0.5)) * f(1.0 / 16777216.0)
    u3 = u[3].astype(f) * inv32
    r0, r1 = np.sqrt(f(-2.0) * np.log(u0)).astype(f), np.sqrt(f(-2.0) * np.log(u2)).astype(f)
    a0, a1 = (f(2.0) * u1).astype(np.float64) * np.pi, (f(2.0) * u3).astype(np.float64) * np.pi
    return (r0 * np.cos(a0).astype(f), r0 * np.sin(a0).astype(f), r1 * np.cos(a1).astype(f), r1 * np.sin(a1).astype(f))


def sample_rotations_np(n: int, seed: int = 0, first_index: int = 0) -> np.ndarray:
    """Restatement of ahv_so3_sample: hypothesis i = quaternion map of the normals Philox(seed, first_index+i)."""
    q = _philox_quaternions(np.arange(first_index, first_index + n, dtype=np.uint64), seed, 0x3d41, 0x4856)
    return rotations_from_normals_np(np.stack(q, axis=-1))


def perturb_rotations_np(R_center: np.ndarray, m: int, max_angle_deg: float, seed: int = 0) -> np.ndarray:
    """Restatement of ahv_so3_perturb: [...,3,3] -> [...,m,3,3]; index 0 = the centre, index j a Haar rotation
    (Philox counter i*m+j) with its angle rescaled from [0,pi] to [0,max_angle] about the same axis, applied on
    the left of the centre."""
    f = np.float32
    c = np.ascontiguousarray(R_center, dtype=f).reshape(-1, 3, 3)
    n = c.shape[0]
    t = np.arange(n * m, dtype=np.uint64)
    qa, qb, qc, qd = _philox_quaternions(t, seed, 0x7065, 0x7274)
    vn = np.sqrt(qb * qb + qc * qc + qd * qd).astype(f)
    theta = (f(2.0) * np.arctan2(vn, np.abs(qa))).astype(f)
    phi = theta * f(np.deg2rad(max_angle_deg) / np.pi)
    sh, ch = np.sin(f(0.5) * phi).astype(f), np.cos(f(0.5) * phi).astype(f)
    k = np.where(vn > 0, sh / np.maximum(vn, f(1e-30)), f(0)).astype(f)
    dR = rotations_from_normals_np(np.stack([ch, k * qb, k * qc, k * qd], axis=-1)).reshape(n, m, 3, 3)
    out = np.einsum("nmij,njk->nmik", dR.astype(np.float64), c.astype(np.float64)).astype(f)
    out[:, 0] = c
    return out.reshape(*np.shape(R_center)[:-2], m, 3, 3)


# --------------------------------------------------------------------------
# base coordinates of F.affine_grid(align_corners=False) for size 8
# ATen builds linspace(-1,1,8)*(7/8); two entries are not exactly (2i+1)/8-1.
# --------------------------------------------------------------------------
def base_coords_np() -> np.ndarray:
    """The 8 base coordinates exactly as ATen produces them on CPU (fixture
    `tests/golden/weights.npz['base']` holds the same values, taken from
    F.affine_grid(eye(3,4), (1,1,8,8,8), align_corners=False))."""
    hexes = ("-0x1.cp-1", "-0x1.4p-1", "-0x1.7ffffep-2", "-0x1.fffff8p-4",
             "0x1.fffff8p-4", "0x1.7ffffep-2", "0x1.4p-1", "0x1.cp-1")
    return np.array([float.fromhex(h) for h in hexes], dtype=np.float32)


# --------------------------------------------------------------------------
# a2: utils.rotate_volume  (utils.py:113-131)  — explicit arithmetic
# --------------------------------------------------------------------------
def rotate_volume_np(vol: np.ndarray, R: np.ndarray, base: np.ndarray | None = None) -> np.ndarray:
    """vol [16,8,8,8] fp32, R [n,3,3] fp32 -> [n,16,8,8,8] fp32.

    grid[n,d,h,w] = R[n] @ (x_w, y_h, z_d)           (F.affine_grid, utils.py:126)
    ix = ((gx+1)*8-1)/2, iy <- gy (H), iz <- gz (D)  (GridSampler.h unnormalize)
    8 taps, weight prod(1-|i-corner|), every tap dropped on its own when
    outside [0,7] (padding_mode='zeros', utils.py:129).
    """
    f = np.float32
    base = base_coords_np() if base is None else base.astype(f)
    n = R.shape[0]
    zz, yy, xx = np.meshgrid(base, base, base, indexing="ij")  # [d,h,w]
    xs, ys, zs = xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)
    R = R.astype(f)
    out = np.zeros((n, C, VOX), dtype=f)
    volf = vol.reshape(C, VOX).astype(f)
    for a in range(0, n, 2048):
        Rb = R[a : a + 2048]
        g = (Rb[:, :, 0:1] * xs[None, None] + Rb[:, :, 1:2] * ys[None, None]).astype(f)
        g = (g + Rb[:, :, 2:3] * zs[None, None]).astype(f)         # [nb,3,512]
        idx = (((g + f(1.0)) * f(8.0) - f(1.0)) / f(2.0)).astype(f)
        i0 = np.floor(idx)
        fr = (idx - i0).astype(f)
        i0 = i0.astype(np.int64)
        acc = np.zeros((Rb.shape[0], C, VOX), dtype=f)
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    x = i0[:, 0] + dx
                    y = i0[:, 1] + dy
                    z = i0[:, 2] + dz
                    wx = fr[:, 0] if dx else (f(1.0) - fr[:, 0])
                    wy = fr[:, 1] if dy else (f(1.0) - fr[:, 1])
                    wz = fr[:, 2] if dz else (f(1.0) - fr[:, 2])
                    wgt = (wx * wy * wz).astype(f)
                    ok = (x >= 0) & (x < W) & (y >= 0) & (y < H) & (z >= 0) & (z < D)
                    lin_i = np.where(ok, (z * H + y) * W + x, 0)
                    vals = volf[:, lin_i]                      # [C,nb,512]
                    acc += np.where(ok, wgt, f(0))[:, None, :] * vals.transpose(1, 0, 2)
        out[a : a + 2048] = acc
    return out.reshape(n, C, D, H, W)


# --------------------------------------------------------------------------
# a4: Feature_Aligner.forward_3d2d  (modules/modules.py:112-124, weights :66-70)
# --------------------------------------------------------------------------
def triplane_np(v: np.ndarray) -> np.ndarray:
    """[m,16,8,8,8] -> [m,384,8,8]; channel order x|(c w), y|(c h), z|(c d)
    (modules/modules.py:115-118)."""
    m = v.shape[0]
    x = v.transpose(0, 1, 4, 2, 3).reshape(m, C * W, D, H)   # b (c w) d h
    y = v.transpose(0, 1, 3, 2, 4).reshape(m, C * H, D, W)   # b (c h) d w
    z = v.reshape(m, C * D, H, W)                            # b (c d) h w
    return np.concatenate([x, y, z], axis=1)


def forward_3d2d_np(v: np.ndarray, W1: np.ndarray, W2: np.ndarray, b2: np.ndarray) -> np.ndarray:
    """[m,16,8,8,8] -> [m,32,64] unit vectors over the channel axis.

    conv1x1 384->32 (no bias), ReLU, conv1x1 32->32 + bias
    (modules/modules.py:66-70), F.normalize(p=2, dim=1) = v/max(||v||,1e-12)
    (:122), flatten(2) (:122)."""
    f = np.float32
    t = triplane_np(v.astype(f)).reshape(v.shape[0], KTRI, NPOS)
    h1 = np.einsum("ok,mkp->mop", W1.reshape(OCH, KTRI).astype(f), t, optimize=True).astype(f)
    h1 = np.maximum(h1, f(0))
    h2 = np.einsum("oi,mip->mop", W2.reshape(OCH, OCH).astype(f), h1, optimize=True).astype(f)
    h2 = (h2 + b2.astype(f)[None, :, None]).astype(f)
    nrm = np.sqrt((h2 * h2).sum(axis=1, keepdims=True)).astype(f)
    return (h2 / np.maximum(nrm, f(1e-12))).astype(f)


# --------------------------------------------------------------------------
# a5/a6: correlate, reduce, select  (modules/model.py:193-196)
# --------------------------------------------------------------------------
def score_np(vol_src, vol_tgt, R, W1, W2, b2, chunk: int = 1024) -> np.ndarray:
    """vol_src/vol_tgt [B,16,8,8,8]; R [N,3,3] shared or [B,N,3,3] per pair
    -> scores [B,N] fp32 = mean over 64 positions of the cosine similarity."""
    f = np.float32
    B = vol_src.shape[0]
    per_pair = R.ndim == 4
    N = R.shape[1] if per_pair else R.shape[0]
    tgt = forward_3d2d_np(vol_tgt, W1, W2, b2)                # [B,32,64]
    out = np.zeros((B, N), dtype=f)
    for b in range(B):
        Rb = R[b] if per_pair else R
        for a in range(0, N, chunk):
            rot = rotate_volume_np(vol_src[b], Rb[a : a + chunk])
            fr = forward_3d2d_np(rot, W1, W2, b2)             # [n,32,64]
            out[b, a : a + chunk] = (fr * tgt[b][None]).sum(axis=1).mean(axis=-1)
    return out


def select_np(scores: np.ndarray, k: int = 1):
    """top-k per pair, descending score, ties -> lowest hypothesis index
    (torch.max on CPU returns the first maximal index, modules/model.py:195)."""
    B, N = scores.shape
    k = min(k, N)
    idx = np.lexsort((np.broadcast_to(np.arange(N), (B, N)), -scores.astype(np.float64)), axis=-1)[:, :k]
    val = np.take_along_axis(scores, idx, axis=1)
    return val, idx.astype(np.int64)


# --------------------------------------------------------------------------
# torch restatement of the idiom: same ATen calls as the reference
# --------------------------------------------------------------------------
def rotate_volume_torch(volume, rotation_matrix, padding_mode="zeros"):
    """utils.py:113-131 restated: [R|0] -> affine_grid -> grid_sample."""
    import torch
    import torch.nn.functional as F

    theta = torch.cat([rotation_matrix, rotation_matrix.new_zeros(rotation_matrix.size(0), 3, 1)], dim=-1)
    grid = F.affine_grid(theta, list(volume.size()), align_corners=False)
    return F.grid_sample(volume, grid, padding_mode=padding_mode, align_corners=False)


def forward_3d2d_torch(v, W1, W2, b2):
    """modules/modules.py:112-124 restated with torch ops (permute == rearrange)."""
    import torch
    import torch.nn.functional as F

    m = v.shape[0]
    x = v.permute(0, 1, 4, 2, 3).reshape(m, C * W, D, H)
    y = v.permute(0, 1, 3, 2, 4).reshape(m, C * H, D, W)
    z = v.reshape(m, C * D, H, W)
    t = torch.cat([x, y, z], dim=1)
    t = F.conv2d(t, W1.reshape(OCH, KTRI, 1, 1))
    t = F.relu(t)
    t = F.conv2d(t, W2.reshape(OCH, OCH, 1, 1), b2)
    return F.normalize(t, p=2, dim=1).flatten(2)


def score_torch(vol_src, vol_tgt, R, W1, W2, b2, chunk: int = 5000):
    """modules/model.py:186-193 restated, chunked over hypotheses (results are
    independent per hypothesis).  R [N,3,3] shared or [B,N,3,3] per pair."""
    import torch

    with torch.no_grad():
        B = vol_src.shape[0]
        per_pair = R.dim() == 4
        N = R.shape[1] if per_pair else R.shape[0]
        tgt = forward_3d2d_torch(vol_tgt, W1, W2, b2)         # [B,32,64]
        out = torch.empty(B, N, dtype=torch.float32, device=vol_src.device)
        for b in range(B):
            Rb = R[b] if per_pair else R
            for a in range(0, N, chunk):
                Rc = Rb[a : a + chunk]
                rot = rotate_volume_torch(vol_src[b][None].expand(Rc.shape[0], -1, -1, -1, -1), Rc)
                fr = forward_3d2d_torch(rot, W1, W2, b2)
                out[b, a : a + chunk] = (fr * tgt[b][None]).sum(dim=1).mean(dim=-1)
        return out


def geodesic_deg_np(Ra: np.ndarray, Rb: np.ndarray) -> np.ndarray:
    """modules/model.py:198-200: arccos((tr(Ra^T Rb)-1)/2) in degrees."""
    s = ((Ra.reshape(-1, 9) * Rb.reshape(-1, 9)).sum(-1).clip(-1, 3) - 1) / 2
    return np.arccos(s) * 180.0 / math.pi


# --------------------------------------------------------------------------
# ctypes binding of the plain-C restatement (oracle/ahv_oracle.c)
# --------------------------------------------------------------------------
_C_LIB = None


def c_lib(build: bool = True):
    """Load oracle/libahv_oracle.so (built by `make -C oracle`)."""
    global _C_LIB
    if _C_LIB is not None:
        return _C_LIB
    import ctypes
    import os
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "libahv_oracle.so")
    src = os.path.join(here, "ahv_oracle.c")
    if build and (not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", here, "libahv_oracle.so"], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(so)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.ahv_oracle_so3_from_normals.argtypes = [fp, fp, ctypes.c_int64]
    lib.ahv_oracle_so3_grid.argtypes = [ctypes.c_int64, ctypes.c_int64, fp, ctypes.c_int64]
    lib.ahv_oracle_so3_grid.restype = None
    lib.ahv_oracle_score.argtypes = [fp, fp, fp, ctypes.c_int, fp, fp, fp, fp, ctypes.c_int, ctypes.c_int64, fp, ctypes.c_int]
    lib.ahv_oracle_argmax.argtypes = [fp, ctypes.c_int, ctypes.c_int64, fp, ctypes.POINTER(ctypes.c_int64)]
    dp = ctypes.POINTER(ctypes.c_double)
    lib.ahv_oracle_score_backward.argtypes = [fp, fp, fp, ctypes.c_int, fp, fp, fp, fp, ctypes.c_int, ctypes.c_int64, fp, dp, dp, dp, dp, dp]
    for fn in (lib.ahv_oracle_so3_from_normals, lib.ahv_oracle_score, lib.ahv_oracle_argmax, lib.ahv_oracle_score_backward):
        fn.restype = None
    _C_LIB = lib
    return lib


def _fp(a):
    import ctypes

    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def rotations_from_normals_c(o: np.ndarray) -> np.ndarray:
    o = np.ascontiguousarray(o, dtype=np.float32)
    R = np.empty((o.shape[0], 9), dtype=np.float32)
    c_lib().ahv_oracle_so3_from_normals(_fp(o), _fp(R), o.shape[0])
    return R.reshape(-1, 3, 3)


def grid_rotations_c(n_total: int, first: int = 0, count: int | None = None) -> np.ndarray:
    """Restatement of ahv_so3_grid (super-Fibonacci SO(3) grid; extension beyond the reference)."""
    count = n_total - first if count is None else count
    R = np.empty((count, 9), dtype=np.float32)
    c_lib().ahv_oracle_so3_grid(n_total, first, _fp(R), count)
    return R.reshape(-1, 3, 3)


def score_c(vol_src, vol_tgt, R, W1, W2, b2, base=None, nthreads: int | None = None) -> np.ndarray:
    """C restatement of modules/model.py:186-193 (pthreads over pair x hypothesis)."""
    import os

    nthreads = nthreads or (os.cpu_count() or 1)
    f = np.float32
    vs = np.ascontiguousarray(vol_src, dtype=f)
    vt = np.ascontiguousarray(vol_tgt, dtype=f)
    Rc = np.ascontiguousarray(R, dtype=f)
    per_pair = Rc.ndim == 4
    B = vs.shape[0]
    N = Rc.shape[1] if per_pair else Rc.shape[0]
    base = base_coords_np() if base is None else np.ascontiguousarray(base, dtype=f)
    W1c = np.ascontiguousarray(W1, dtype=f).reshape(OCH, KTRI)
    W2c = np.ascontiguousarray(W2, dtype=f).reshape(OCH, OCH)
    b2c = np.ascontiguousarray(b2, dtype=f)
    out = np.empty((B, N), dtype=f)
    c_lib().ahv_oracle_score(_fp(vs), _fp(vt), _fp(Rc), int(per_pair), _fp(W1c), _fp(W2c), _fp(b2c), _fp(base), B, N, _fp(out), int(nthreads))
    return out


def score_backward_c(vol_src, tgt_feat, R, W1, W2, b2, grad_scores, base=None):
    """C restatement (double accumulation, scatter-form adjoint) of the gradient of the scores with respect
    to the source volumes, the target FEATURES, W1, W2 and b2 (modules/model.py:53-56 under autograd).
    Returns float64 arrays (g_vol [B,16,8,8,8], g_tgt [B,32,64], g_W1 [32,384], g_W2 [32,32], g_b2 [32])."""
    import ctypes

    f = np.float32
    vs = np.ascontiguousarray(vol_src, dtype=f)
    tg = np.ascontiguousarray(tgt_feat, dtype=f)
    Rc = np.ascontiguousarray(R, dtype=f)
    gs = np.ascontiguousarray(grad_scores, dtype=f)
    per_pair = Rc.ndim == 4
    B = vs.shape[0]
    N = Rc.shape[1] if per_pair else Rc.shape[0]
    base = base_coords_np() if base is None else np.ascontiguousarray(base, dtype=f)
    W1c = np.ascontiguousarray(W1, dtype=f).reshape(OCH, KTRI)
    W2c = np.ascontiguousarray(W2, dtype=f).reshape(OCH, OCH)
    b2c = np.ascontiguousarray(b2, dtype=f)
    g_vol, g_tgt = np.zeros((B, 16, 8, 8, 8)), np.zeros((B, OCH, 64))
    g_W1, g_W2, g_b2 = np.zeros((OCH, KTRI)), np.zeros((OCH, OCH)), np.zeros(OCH)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    c_lib().ahv_oracle_score_backward(_fp(vs), _fp(tg), _fp(Rc), int(per_pair), _fp(W1c), _fp(W2c), _fp(b2c), _fp(base), B, N,
                                      _fp(gs), dp(g_vol), dp(g_tgt), dp(g_W1), dp(g_W2), dp(g_b2))
    return g_vol, g_tgt, g_W1, g_W2, g_b2
