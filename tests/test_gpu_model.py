"""GPU: the drop-in Estimator steps run the fused path and agree with the unfused idiom."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _TinyBackbone(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.proj = torch.nn.Conv2d(3, 768, 1)

    def forward(self, img):
        return self.proj(torch.nn.functional.adaptive_avg_pool2d(img, 8))


def _cfg(n=512):
    return {"DATA": {"NUM_ROTA": n, "BG": False, "SIZE_THR": 10, "ACC_THR": 15}}


def _batch(dev, B=3):
    g = torch.Generator().manual_seed(5)
    img = lambda: torch.rand(B, 3, 64, 64, generator=g).to(dev)
    rot = lambda: torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0].to(dev)
    return {"src_img": img(), "ref_img": img(), "src_mask": torch.ones(B, 1, 64, 64, device=dev),
            "ref_mask": torch.ones(B, 1, 64, 64, device=dev), "src_R": rot(), "ref_R": rot()}


def test_test_step_matches_unfused_idiom(ahv):
    from modules.model import Estimator

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = Estimator(_cfg(), feature_extractor=_TinyBackbone()).to(dev).eval()
    batch = _batch(dev)
    torch.manual_seed(11)
    with torch.no_grad():
        m.test_step(batch, 0)
    assert len(m.step_outputs) == 1 and len(m.gt_dis) == 1 and len(m.pred_Rs) == 1
    assert m.step_outputs[0].shape == (3,) and m.pred_Rs[0].shape == (27,)
    # recompute with the reference-shaped unfused ops on the same hypothesis set
    torch.manual_seed(11)
    R = ahv.so3.random_rotations(512, device=dev)
    with torch.no_grad():
        vs, vt = m(batch["src_img"], batch["src_mask"], batch["ref_img"], batch["ref_mask"])
        rot = torch.stack([ahv.refcompat.rotate_volume(v[None].expand(512, -1, -1, -1, -1), R) for v in vs])
        f = m.feature_aligner.forward_3d2d(rot.reshape(-1, 16, 8, 8, 8)).reshape(3, 512, -1, 64)
        t = m.feature_aligner.forward_3d2d(vt)
        sim = (f * t[:, None]).sum(dim=2).mean(dim=-1)
        best, idx = torch.max(sim, dim=1)
    pred = torch.from_numpy(m.pred_Rs[0]).reshape(3, 3, 3).to(dev)
    fused_sim = m.score_rotations(vs, vt, R)
    assert torch.allclose(fused_sim, sim, rtol=1e-3, atol=0)
    picked = sim[torch.arange(3), (R[None] == pred[:, None]).all(-1).all(-1).float().argmax(1)]
    assert torch.all((best - picked).abs() <= 2e-3 * best.abs())     # same top-1 or an equal-score tie
    gt = torch.bmm(batch["ref_R"], torch.inverse(batch["src_R"]))
    s = ((pred.reshape(-1, 9) * gt.reshape(-1, 9)).sum(-1).clamp(-1, 3) - 1) / 2
    assert torch.allclose(m.step_outputs[0], torch.arccos(s) * 180 / math.pi, atol=1e-3)


def test_validation_step_and_gt_hypothesis(ahv):
    from modules.model import Estimator

    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    m = Estimator(_cfg(256), feature_extractor=_TinyBackbone()).to(dev).eval()
    batch = _batch(dev)
    with torch.no_grad():
        m.validation_step(batch, 0)
        vs, vt = m(batch["src_img"], batch["src_mask"], batch["ref_img"], batch["ref_mask"])
        gt = torch.bmm(batch["ref_R"], torch.inverse(batch["src_R"]))
        # modules/model.py:137-143 unfused: rotate each pair's own volume by its own GT rotation
        rot = ahv.refcompat.rotate_volume(vs, gt)
        gt_sim = (m.feature_aligner.forward_3d2d(rot) * m.feature_aligner.forward_3d2d(vt)).sum(dim=1).mean(dim=-1)
    assert m.last_gt_sim.shape == (3,)
    assert torch.allclose(m.last_gt_sim, gt_sim, rtol=1e-3, atol=0)
    assert len(m.step_outputs) == 1
    m.on_validation_epoch_end()
    assert m.step_outputs == []
    small = dict(batch)
    small["src_mask"] = torch.zeros_like(batch["src_mask"])
    assert m.test_step(small, 0) == 0 and len(m.pred_Rs) == 0        # "Skip bad case" (modules/model.py:173-175)


def test_co3d_estimator_with_named_backbone(ahv):
    from modules.model_co3d import Estimator

    dev = torch.device("cuda", 0)
    torch.manual_seed(2)
    m = Estimator(_cfg(3000)).to(dev).eval()            # SwinV2-T stand-in, random init
    img = torch.randn(2, 3, 256, 256, device=dev)
    gt = torch.eye(3, device=dev).expand(2, 3, 3).contiguous()
    torch.manual_seed(3)
    proposals = ahv.so3.random_rotations(3000, device=dev)
    err = m.evaluate_pairs(img, img.flip(0), gt, proposals)
    assert err.shape == (2,) and torch.isfinite(err).all() and float(err.max()) <= 180.0
    val, idx, Rb, _ = m.predict(img, img.flip(0), proposals, k=4)
    assert idx.shape == (2, 4) and torch.equal(Rb[:, 0], proposals[idx[:, 0]])


def test_refine_improves_or_keeps_score(ahv, golden):
    dev = torch.device("cuda", 0)
    g, w = golden["shared_n3000_b3"], golden["weights"]
    T = lambda a: torch.from_numpy(a).to(dev)
    v = ahv.HypothesisVerifier(T(w["W1"]), T(w["W2"]), T(w["b2"]))
    Rb, val, first, cand = v.refine(T(g["vol_src"]), T(g["vol_tgt"]), T(g["R"]), k=8, m=32, max_angle_deg=4.0)
    # index 0 of every perturbation set is the centre itself, so pass 2 can never score lower than pass 1
    assert torch.all(val >= first.topk_val[:, 0] - 1e-6)
    assert cand.shape == (3, 8 * 32, 3, 3)
    assert torch.allclose(Rb @ Rb.transpose(1, 2), torch.eye(3, device=dev).expand(3, 3, 3), atol=1e-5)


def test_training_step_backpropagates_to_backbone_and_head(ahv):
    """modules/model.py:78-116 / model_co3d.py:71-96: one optimisation step lowers the InfoNCE loss."""
    from modules.model import Estimator
    from modules.model_co3d import Estimator as EstimatorCo3d

    dev = torch.device("cuda", 0)
    cfg = _cfg(64)
    cfg["TRAIN"] = {"MASK": True, "MASK_RATIO": 0.25, "LR": 1e-3}
    torch.manual_seed(0)
    m = Estimator(cfg, feature_extractor=_TinyBackbone()).to(dev).train()
    batch = _batch(dev)
    opt = m.configure_optimizers()[0][0]
    torch.manual_seed(1)
    loss0 = m.training_step(batch, 0)
    assert loss0.requires_grad and torch.isfinite(loss0)
    opt.zero_grad()
    loss0.backward()
    head = m.feature_aligner.feature_embedding_2d
    for p in (head[0].weight, head[2].weight, head[2].bias, m.feature_extractor.proj.weight,
              m.feature_aligner.feature_embedding_3d.conv1.weight):
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0
    assert m.Rs.shape == (3, 64, 3, 3)
    # CO3D flavour
    c = EstimatorCo3d(cfg, feature_extractor=_TinyBackbone()).to(dev).train()
    gt = torch.linalg.qr(torch.randn(3, 3, 3))[0].to(dev)
    b2 = {"image": torch.stack([batch["src_img"], batch["ref_img"]], 1), "relative_rotation": gt[:, None]}
    loss = c.training_step(b2, 0)
    loss.backward()
    assert torch.isfinite(loss) and c.feature_aligner.feature_embedding_2d[0].weight.grad.abs().sum() > 0
    assert len(c.configure_optimizers()[0][0].param_groups) == 2
    # cfg["TRAIN"]["FAST_BACKWARD"]: same step through the saved-activation tensor-core backward - same loss (the forward
    # arithmetic is the same kernel's), head gradients within the mode's stated accuracy
    grads_exact = [p.grad.clone() for p in (head[0].weight, head[2].weight, head[2].bias)]
    m.fast_backward = True
    opt.zero_grad()
    torch.manual_seed(1)
    loss1 = m.training_step(batch, 0)
    loss1.backward()
    assert torch.allclose(loss1.detach(), loss0.detach(), rtol=1e-5, atol=1e-6)
    for p, g0 in zip((head[0].weight, head[2].weight, head[2].bias), grads_exact):
        assert float((p.grad - g0).norm() / g0.norm()) <= 5e-2
