"""Generate tests/golden/training_grads.npz by running the REFERENCE's own code under autograd on CPU.

Test infrastructure only; runs only in the build container (needs /root/reference).  It does not touch the
other fixtures.

    python oracle/make_golden_training.py

The training path of the reference is `Estimator.infoNCE_loss` (modules/model.py:43-63): per-pair rotation
sets, `rotate_volume` (utils.py:113-131) -> `Feature_Aligner.forward_3d2d` (modules/modules.py:112-124) ->
similarity -> softmax over hypotheses with temperature 0.1.  `Estimator` itself cannot be imported
(lightning / pytorch3d / timm absent), so the body of `infoNCE_loss` is executed here line by line on the
reference's imported `rotate_volume` and `Feature_Aligner` (ACC_THR = 15, config.yaml), and PyTorch autograd
gives the gradients with respect to the two volumes and the head weights - exactly what the reference's
`loss.backward()` produces upstream of the hot path.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ahv_oracle as orc  # noqa: E402
from oracle.make_golden import import_reference  # noqa: E402

ACC_THR = 15.0  # config.yaml DATA.ACC_THR


def infonce_reference(rotate_volume, fa, img_feat_1, img_feat_2, sampled_R, gt_delta_R):
    """modules/model.py:43-63, verbatim in structure; returns (loss [bs], sim [bs, n])."""
    bs = gt_delta_R.shape[0]
    num_rota = sampled_R.shape[1]
    with torch.no_grad():
        gt_sim = (torch.sum(sampled_R.flatten(2) * gt_delta_R.view(-1, 1, 9), dim=-1).clamp(-1, 3) - 1) / 2
        gt_dis = torch.arccos(gt_sim) / np.pi
        posi_indices = [torch.nonzero(180 * gt_dis[i] <= ACC_THR).squeeze(-1) for i in range(bs)]
    img_feat_warp = [rotate_volume(img_feat_1[idx:idx + 1].expand(num_rota, -1, -1, -1, -1), sampled_R[idx]) for idx in range(bs)]
    img_feat_warp = [fa.forward_3d2d(img_feat) for img_feat in img_feat_warp]
    img_feat_2 = fa.forward_3d2d(img_feat_2)
    sim = [(img_feat_warp[idx] * img_feat_2[idx:idx + 1]).sum(dim=1).mean(dim=-1) for idx in range(bs)]
    positive_sim = torch.stack([torch.exp(sim[idx][posi_indices[idx]] / 0.1).sum(dim=0) for idx in range(bs)])
    positive_negative_sim = (torch.exp(torch.stack(sim) / 0.1)).sum(dim=-1)
    loss = -torch.log(positive_sim / positive_negative_sim.clamp(min=1e-8))
    return loss, torch.stack(sim)


def main():
    rotate_volume, Feature_Aligner = import_reference()
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "shared_n3000_b3.npz")))
    w = dict(np.load(os.path.join(ROOT, "tests", "golden", "weights.npz")))
    torch.manual_seed(0)
    fa = Feature_Aligner(in_channel=768, mid_channel=256, out_channel=32, n_heads=4, depth=4).double()
    head = fa.feature_embedding_2d
    with torch.no_grad():  # the committed head weights (same as every other fixture)
        head[0].weight.copy_(torch.from_numpy(w["W1"]).double().reshape(32, 384, 1, 1))
        head[2].weight.copy_(torch.from_numpy(w["W2"]).double().reshape(32, 32, 1, 1))
        head[2].bias.copy_(torch.from_numpy(w["b2"]).double())
    bs, n = 3, 96
    vol_src = torch.from_numpy(gold["vol_src"]).double().requires_grad_(True)
    vol_tgt = torch.from_numpy(gold["vol_tgt"]).double().requires_grad_(True)
    g = torch.Generator().manual_seed(4321)
    gt = orc.rotations_from_normals_torch(torch.randn((bs, 4), generator=g))
    # hypotheses: the ground truth first (modules/model.py:103), a few near it (positives), the rest random
    near = []
    for b in range(bs):
        axis_angle = torch.randn((5, 3), generator=g)
        axis_angle = axis_angle / axis_angle.norm(dim=1, keepdim=True) * (torch.rand((5, 1), generator=g) * 0.2)
        K = torch.zeros(5, 3, 3)
        K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -axis_angle[:, 2], axis_angle[:, 1], axis_angle[:, 2]
        K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -axis_angle[:, 0], -axis_angle[:, 1], axis_angle[:, 0]
        near.append(torch.matrix_exp(K) @ gt[b])
    rnd = orc.rotations_from_normals_torch(torch.randn((bs * (n - 6), 4), generator=g)).reshape(bs, n - 6, 3, 3)
    sampled_R = torch.cat([gt[:, None], torch.stack(near), rnd], dim=1).contiguous()  # [bs, n, 3, 3] fp32
    loss, sim = infonce_reference(rotate_volume, fa, vol_src, vol_tgt, sampled_R.double(), gt.double())
    loss.mean().backward()
    out = os.path.join(ROOT, "tests", "golden", "training_grads.npz")
    np.savez_compressed(
        out,
        sampled_R=sampled_R.numpy(), gt_R=gt.numpy(), acc_thr=np.float64(ACC_THR),
        sim=sim.detach().numpy(), loss=loss.detach().numpy(),
        # gradients of loss.mean(), computed in float64, stored as float32 (the gates are 2e-4 of the maximum)
        g_vol_src=vol_src.grad.float().numpy(), g_vol_tgt=vol_tgt.grad.float().numpy(),
        g_W1=head[0].weight.grad.reshape(32, 384).float().numpy(), g_W2=head[2].weight.grad.reshape(32, 32).float().numpy(),
        g_b2=head[2].bias.grad.float().numpy(),
    )
    print(out, os.path.getsize(out), "loss", loss.detach().numpy())


if __name__ == "__main__":
    main()
