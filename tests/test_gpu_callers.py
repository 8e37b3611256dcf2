"""GPU: the reference's evaluation loops as callers of the fused path, checked against the
ORACLE - the reference's own ATen idiom (oracle.score_torch: F.affine_grid / F.grid_sample / 1x1 convs /
normalize, modules/model.py:186-196) executed on the CPU on the very volumes the fused path was given.

    3dahv_b200.evaluate.evaluate_category / evaluate_pairwise   <->  test_co3d.py:93-198
    3dahv_b200.evaluate.test_category                           <->  test_linemod.py:20-84
    Estimator.test_step / validation_step                       <->  modules/model.py:118-209
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
TOL = 1e-3


class _TinyBackbone(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.proj = torch.nn.Conv2d(3, 768, 1)

    def forward(self, img):
        return self.proj(torch.nn.functional.adaptive_avg_pool2d(img, 8))


class _SyntheticCo3d:
    """Duck-typed like data_loader_co3d.Co3dDataset as test_co3d.py uses it (:109-113)."""

    def __init__(self, n_seq, n_frames, seed):
        g = torch.Generator().manual_seed(seed)
        self.images = [torch.rand(n_frames, 3, 64, 64, generator=g) for _ in range(n_seq)]
        self.R = [torch.linalg.qr(torch.randn(n_frames, 3, 3, generator=g))[0] for _ in range(n_seq)]
        self.n_frames = n_frames

    def __iter__(self):
        for i in range(len(self.images)):
            yield {"n": self.n_frames, "model_id": f"seq{i}"}

    def get_data(self, sequence_name, ids):
        i = int(sequence_name[3:])
        ids = torch.as_tensor(np.asarray(ids))
        return {"image": self.images[i][ids], "R": self.R[i][ids]}


def _cfg(n):
    return {"DATA": {"NUM_ROTA": n, "BG": False, "SIZE_THR": 10, "ACC_THR": 15}}


def _head(model):
    h = model.feature_aligner.feature_embedding_2d
    return (h[0].weight.detach().reshape(32, 384).cpu(), h[2].weight.detach().reshape(32, 32).cpu(), h[2].bias.detach().cpu())


def _geo(Ra, Rb):
    s = ((Ra.reshape(-1, 9) * Rb.reshape(-1, 9)).sum(-1).clamp(-1, 3) - 1) / 2
    return torch.arccos(s) * 180.0 / math.pi


class _Handle:
    def __init__(self, model):
        self.model = model

    def remove(self):
        del self.model.forward          # back to the class's method


def _record_volumes(model):
    """Record the (vol_src, vol_tgt) every `forward` returns - called as model(...) or as self.forward(...)."""
    rec, orig = [], model.forward

    def forward(*args, **kwargs):
        out = orig(*args, **kwargs)
        rec.append((out[0].detach().cpu(), out[1].detach().cpu()))
        return out

    model.forward = forward
    return rec, _Handle(model)


def _check_against_oracle(oracle, sim_ref, picked, err, gt, proposals_cpu):
    """picked [P] indices chosen by the fused path; sim_ref [P,N] the oracle's pred_sim for the same pairs."""
    best, arg = sim_ref.max(dim=1)
    chosen = sim_ref[torch.arange(sim_ref.shape[0]), torch.from_numpy(picked)]
    same = torch.from_numpy(picked) == arg
    assert torch.all(same | ((best - chosen).abs() <= 2 * TOL * best.abs())), "top-1 differs beyond an equal-score tie"
    # where the index agrees the angular error is the oracle's
    err_ref = _geo(proposals_cpu[arg], gt)
    assert np.allclose(err[same.numpy()], err_ref[same].numpy(), atol=2e-3)
    # and in every case it is the error of the rotation that was picked
    assert np.allclose(err, _geo(proposals_cpu[torch.from_numpy(picked)], gt).numpy(), atol=2e-3)


def test_evaluate_category_matches_reference_loop(ahv, oracle):
    """test_co3d.py:93-154: proposals once per category, NP2 ordered frame pairs per sequence, arg-max
    hypothesis, geodesic error - batched + fused here, replayed pair by pair with the oracle."""
    from modules.model_co3d import Estimator

    torch.manual_seed(0)
    model = Estimator(_cfg(2000), feature_extractor=_TinyBackbone()).to(DEV).eval()
    ds = _SyntheticCo3d(n_seq=3, n_frames=5, seed=1)
    rec, handle = _record_volumes(model)
    np.random.seed(0)
    torch.manual_seed(0)                                           # test_co3d.py:24-25
    err, picked, proposals = ahv.evaluate.evaluate_category(_cfg(2000), model, ds, num_frames=3, device=DEV, batch_pairs=8,
                                                            return_details=True)
    handle.remove()
    perms = ahv.evaluate.get_permutations(3)
    assert perms.tolist() == [[0, 1], [0, 2], [1, 0], [1, 2], [2, 0], [2, 1]]
    assert err.shape == (3 * 6,) and proposals.shape == (2000, 3, 3)
    torch.manual_seed(0)
    assert torch.equal(proposals, ahv.so3.random_rotations(2000, device=DEV))   # the per-category draw
    # replay of the reference loop's bookkeeping (same np.random draws, same pair order) for the ground truth
    np.random.seed(0)
    gts = []
    for md in ds:
        kf = np.random.choice(md["n"], 3, replace=False)
        Rp = ds.get_data(md["model_id"], kf)["R"][perms]
        gts.append(torch.bmm(Rp[:, 0].transpose(1, 2), Rp[:, 1]))
    gt = torch.cat(gts)
    vs, vt = torch.cat([r[0] for r in rec]), torch.cat([r[1] for r in rec])   # the volumes the fused path saw
    assert vs.shape[0] == 18
    W1, W2, b2 = _head(model)
    sim_ref = oracle.score_torch(vs, vt, proposals.cpu(), W1, W2, b2)
    _check_against_oracle(oracle, sim_ref, picked, err, gt, proposals.cpu())
    # category statistics as test_co3d.py:181-182 derives them
    errors, e30, e15 = ahv.evaluate.evaluate_pairwise(cfg=_cfg(500), model=model, categories=("a", "b"), num_frames=2,
                                                      print_results=False, device=DEV,
                                                      get_dataset=lambda cfg, category, split, dataset: _SyntheticCo3d(2, 4, ord(category)))
    assert set(errors) == {"a", "b", "mean"} and 0 <= e15["mean"] <= e30["mean"] <= 100
    assert np.isclose(errors["mean"], np.mean([errors["a"], errors["b"]]))


def _linemod_batches(n_batches, B, seed, bad=()):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n_batches):
        rot = lambda: torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0]
        mask = torch.zeros(B, 1, 64, 64) if i in bad else torch.ones(B, 1, 64, 64)
        out.append({"src_img": torch.rand(B, 3, 64, 64, generator=g), "ref_img": torch.rand(B, 3, 64, 64, generator=g),
                    "src_mask": mask, "ref_mask": torch.ones(B, 1, 64, 64), "src_R": rot(), "ref_R": rot()})
    return out


def test_linemod_test_category_matches_reference_loop(ahv, oracle):
    """test_linemod.py:20-84: a fresh codebook per batch, skip of small masks, codebook + ground-truth hypothesis
    scored, error / Acc@30 / Acc@15 / pred_Rs."""
    from modules.model import Estimator

    torch.manual_seed(0)
    cfg = _cfg(1500)
    model = Estimator(cfg, feature_extractor=_TinyBackbone()).to(DEV).eval()
    batches = _linemod_batches(4, 2, seed=3, bad=(1,))
    rec, handle = _record_volumes(model)
    torch.manual_seed(5)
    mean_err, acc30, acc15, pred_Rs, errs, gt_sims = ahv.evaluate.test_category(cfg, model, batches, device=DEV, return_details=True)
    handle.remove()
    assert pred_Rs.shape == (6, 9) and errs.shape == (6,) and len(rec) == 3          # batch 1 skipped (:39-41)
    assert np.isclose(mean_err, errs.mean(), atol=1e-4)
    assert np.isclose(acc30, 100 * np.mean(errs < 30)) and np.isclose(acc15, 100 * np.mean(errs < 15))
    W1, W2, b2 = _head(model)
    torch.manual_seed(5)
    good = [b for i, b in enumerate(batches) if i != 1]
    for j, (data, (vs, vt)) in enumerate(zip(good, rec)):
        codebook = ahv.so3.random_rotations(1500, device=DEV).cpu()                 # the draw test_category made for this batch
        gt = torch.bmm(data["ref_R"], torch.inverse(data["src_R"]))
        sim_ref = oracle.score_torch(vs, vt, codebook, W1, W2, b2)
        picked = np.array([int((codebook == torch.from_numpy(pred_Rs[2 * j + b]).reshape(3, 3)).all(-1).all(-1).float().argmax())
                           for b in range(2)])
        _check_against_oracle(oracle, sim_ref, picked, errs[2 * j:2 * j + 2], gt, codebook)
        # gt_sim (test_linemod.py:53-60): per-pair rotation, N = 1
        gt_ref = oracle.score_torch(vs, vt, gt[:, None].contiguous(), W1, W2, b2)[:, 0]
        assert np.allclose(gt_sims[2 * j:2 * j + 2], gt_ref.numpy(), rtol=TOL, atol=0)


def test_estimator_steps_vs_oracle_idiom(ahv, oracle):
    """Estimator.test_step / validation_step (modules/model.py:118-209) against the reference idiom on the CPU."""
    from modules.model import Estimator

    torch.manual_seed(0)
    model = Estimator(_cfg(1024), feature_extractor=_TinyBackbone()).to(DEV).eval()
    batch = {k: v.to(DEV) for k, v in _linemod_batches(1, 3, seed=8)[0].items()}
    rec, handle = _record_volumes(model)
    torch.manual_seed(11)
    with torch.no_grad():
        model.test_step(batch, 0)
    handle.remove()
    torch.manual_seed(11)
    R = ahv.so3.random_rotations(1024, device=DEV).cpu()                             # modules/model.py:184
    vs, vt = rec[0]
    W1, W2, b2 = _head(model)
    sim_ref = oracle.score_torch(vs, vt, R, W1, W2, b2)                             # :186-193 on the CPU
    pred = torch.from_numpy(model.pred_Rs[0]).reshape(3, 3, 3)
    picked = np.array([int((R == pred[b]).all(-1).all(-1).float().argmax()) for b in range(3)])
    gt = torch.bmm(batch["ref_R"], torch.inverse(batch["src_R"])).cpu()
    _check_against_oracle(oracle, sim_ref, picked, model.step_outputs[0].cpu().numpy(), gt, R)
    assert np.allclose(model.gt_dis[0].cpu().numpy(), _geo(batch["src_R"].cpu(), batch["ref_R"].cpu()).numpy(), atol=2e-3)
    # validation_step: the ground-truth hypothesis (:137-143)
    with torch.no_grad():
        model.validation_step(batch, 0)
    gt_ref = oracle.score_torch(vs, vt, gt[:, None].contiguous(), W1, W2, b2)[:, 0]
    assert np.allclose(model.last_gt_sim.cpu().numpy(), gt_ref.numpy(), rtol=TOL, atol=0)
    # the fused scores themselves
    fused = model.score_rotations(vs.to(DEV), vt.to(DEV), R.to(DEV)).cpu()
    assert float(((fused - sim_ref).abs() / sim_ref.abs().clamp_min(1e-6)).max()) <= TOL


def test_objaverse_driver(ahv, tmp_path):
    """test_objaverse.py:14-46: Trainer.test -> test_step per batch, then error / Acc@30 / Acc@15 and the files."""
    from modules.model import Estimator

    torch.manual_seed(0)
    cfg = _cfg(800)
    model = Estimator(cfg, feature_extractor=_TinyBackbone()).to(DEV).eval()
    batches = _linemod_batches(3, 2, seed=4, bad=(2,))
    model.step_outputs.append(torch.zeros(1, device=DEV))               # stale entries are cleared like the script does
    err, acc30, acc15, pred_Rs = ahv.evaluate.test_objaverse(cfg, model, batches, device=DEV, out_dir=str(tmp_path))
    assert pred_Rs.shape == (4, 9) and len(model.step_outputs) == 2 and len(model.gt_dis) == 2
    errs = torch.cat(model.step_outputs).cpu().numpy()
    assert np.isclose(err, errs.mean(), atol=1e-4) and np.isclose(acc30, 100 * np.mean(errs < 30))
    R = torch.from_numpy(pred_Rs).reshape(4, 3, 3).float()
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(4, 3, 3), atol=1e-5)
    saved = np.loadtxt(tmp_path / "objaverse_pred_Rs.txt")
    assert np.allclose(saved, pred_Rs) and "avg_acc_30" in (tmp_path / "result.txt").read_text()
