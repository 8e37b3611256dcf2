"""Generate tests/golden/*.npz by running the REFERENCE's own code on CPU.

Runs only in the build container (needs /root/reference); the fixtures it
writes are committed so the GPU box never needs the reference tree.

    python oracle/make_golden.py            # rewrites tests/golden/

What is imported from the reference (unmodified, via the stub recipe of
SURVEY.md Appendix A): `utils.rotate_volume` (utils.py:113-131) and
`modules.modules.Feature_Aligner` (modules/modules.py:49-124).  The scoring
idiom itself is not a function in the reference; it is executed here exactly
as written at modules/model.py:186-196 (and :137-143 for per-pair rotations).
`random_rotations` is pytorch3d (absent): the rotation sets are produced by the
restated sampler and STORED in the fixture, so score parity does not depend on
it.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ahv_oracle as orc  # noqa: E402


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "pytorch3d", "pytorch3d.transforms"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pytorch3d.transforms"].matrix_to_rotation_6d = lambda x: x
    sys.path.insert(0, "/root/reference")
    from utils import rotate_volume  # utils.py:113
    from modules.modules import Feature_Aligner  # modules/modules.py:49
    import modules.modules as _ref_mod

    # this repo also has modules/ and transformer/ (drop-in API): make sure the REFERENCE's were imported
    assert _ref_mod.__file__.startswith("/root/reference/"), _ref_mod.__file__
    assert sys.modules["transformer.attention"].__file__.startswith("/root/reference/")

    return rotate_volume, Feature_Aligner


def reference_scores(rotate_volume, fa, vol_src, vol_tgt, sampled_R):
    """modules/model.py:186-196 verbatim in structure (shared rotation set)."""
    B, C, D, H, W = vol_src.shape
    n = sampled_R.shape[0]
    feats = [rotate_volume(v[None].expand(n, -1, -1, -1, -1), sampled_R) for v in vol_src]
    feats = torch.stack(feats).reshape(-1, C, D, H, W)
    feats = fa.forward_3d2d(feats).reshape(B, n, -1, H * W)
    tgt = fa.forward_3d2d(vol_tgt)
    pred_sim = (feats * tgt[:, None]).sum(dim=2).mean(dim=-1)
    best, idx = torch.max(pred_sim, dim=1)
    return pred_sim, best, idx, tgt


def reference_scores_per_pair(rotate_volume, fa, vol_src, vol_tgt, R_bn):
    """modules/model.py:53-56 (infoNCE forward) / :137-143: per-pair rotations."""
    B = vol_src.shape[0]
    n = R_bn.shape[1]
    warp = [rotate_volume(vol_src[i : i + 1].expand(n, -1, -1, -1, -1), R_bn[i]) for i in range(B)]
    warp = [fa.forward_3d2d(w) for w in warp]
    tgt = fa.forward_3d2d(vol_tgt)
    sim = [(warp[i] * tgt[i : i + 1]).sum(dim=1).mean(dim=-1) for i in range(B)]
    return torch.stack(sim)


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    rotate_volume, Feature_Aligner = import_reference()
    torch.manual_seed(0)  # mirrors test_co3d.py:24
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    fa = Feature_Aligner(in_channel=768, mid_channel=256, out_channel=32, n_heads=4, depth=4).eval()
    sd = fa.state_dict()
    W1 = sd["feature_embedding_2d.0.weight"].reshape(32, 384).clone()
    W2 = sd["feature_embedding_2d.2.weight"].reshape(32, 32).clone()
    b2 = sd["feature_embedding_2d.2.bias"].clone()

    with torch.no_grad():
        B = 3
        vol_src, vol_tgt = fa.forward_2d3d(
            torch.randn(B, 768, 8, 8), torch.randn(B, 768, 8, 8), random_mask=False, mask_ratio=0.0
        )
        g = torch.Generator().manual_seed(1234)
        normals = torch.randn((3000, 4), generator=g)
        R = orc.rotations_from_normals_torch(normals)

        # -- shared rotation set, config-1 size (N=3000, config.yaml:10) on 3 pairs
        sim, best, idx, tgt = reference_scores(rotate_volume, fa, vol_src, vol_tgt, R)

        # -- rotate_volume / forward_3d2d on their own (small)
        Rs = R[:7]
        rot = rotate_volume(vol_src[0][None].expand(7, -1, -1, -1, -1), Rs)
        f3d = fa.forward_3d2d(rot)

        # -- special rotations: identity, axis flips, 90-degree turns, duplicates (exact ties)
        eye = torch.eye(3)
        rz90 = torch.tensor([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
        rx90 = torch.tensor([[1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, 1.0, 0.0]])
        ry180 = torch.diag(torch.tensor([-1.0, 1.0, -1.0]))
        special = torch.stack([eye, rz90, rx90, ry180, R[5], R[5], eye, R[11]])
        sim_sp, best_sp, idx_sp, _ = reference_scores(rotate_volume, fa, vol_src, vol_tgt, special)
        rot_sp = rotate_volume(vol_src[1][None].expand(8, -1, -1, -1, -1), special)

        # -- per-pair rotations (training / GT-hypothesis shape, N=1 and N=5)
        R_bn = orc.rotations_from_normals_torch(torch.randn((B * 5, 4), generator=g)).reshape(B, 5, 3, 3)
        sim_pp = reference_scores_per_pair(rotate_volume, fa, vol_src, vol_tgt, R_bn)
        sim_gt = reference_scores_per_pair(rotate_volume, fa, vol_src, vol_tgt, R_bn[:, :1])

        # -- base coordinate table as ATen builds it
        base = torch.nn.functional.affine_grid(torch.eye(3, 4)[None], (1, 1, 8, 8, 8), align_corners=False)[0, 0, 0, :, 0]

        # -- a volume with large dynamic range / zeros (edge case: empty volume -> zero norm)
        vol_zero = torch.zeros(1, 16, 8, 8, 8)
        sim_zero, _, idx_zero, _ = reference_scores(rotate_volume, fa, vol_zero, vol_tgt[:1], R[:16])

    np.savez_compressed(
        os.path.join(out_dir, "weights.npz"),
        W1=W1.numpy(), W2=W2.numpy(), b2=b2.numpy(), base=base.numpy(),
    )
    np.savez_compressed(
        os.path.join(out_dir, "shared_n3000_b3.npz"),
        vol_src=vol_src.numpy(), vol_tgt=vol_tgt.numpy(), normals=normals.numpy(), R=R.numpy(),
        scores=sim.numpy(), best=best.numpy(), best_idx=idx.numpy(), tgt_feat=tgt.numpy(),
    )
    np.savez_compressed(
        os.path.join(out_dir, "primitives.npz"),
        R=Rs.numpy(), rotated=rot.numpy(), feat=f3d.numpy(),
        special_R=special.numpy(), special_scores=sim_sp.numpy(), special_idx=idx_sp.numpy(),
        special_rotated=rot_sp.numpy(),
    )
    np.savez_compressed(
        os.path.join(out_dir, "per_pair.npz"),
        R=R_bn.numpy(), scores=sim_pp.numpy(), scores_gt=sim_gt.numpy(),
        zero_scores=sim_zero.numpy(), zero_idx=idx_zero.numpy(),
    )
    for f in sorted(os.listdir(out_dir)):
        print(f, os.path.getsize(os.path.join(out_dir, f)))


if __name__ == "__main__":
    main()
