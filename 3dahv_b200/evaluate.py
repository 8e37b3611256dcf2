"""The reference's evaluation loops as first-class callers of the fused path.

    evaluate_category / evaluate_pairwise   test_co3d.py:93-198     (CO3D: per-category proposals, NP2 frame pairs)
    test_category                           test_linemod.py:20-84   (LINEMOD: DataLoader of pair batches, gt similarity)
    test_objaverse                          test_objaverse.py:14-46 (Objaverse: Trainer.test -> Estimator.test_step)

Same arguments, same random draws in the same order (`random_rotations` from torch's CPU generator once per
category / once per batch, `np.random.choice` for the key frames), same returned statistics.  What changes is
what happens between `model(...)` and the angular error: the reference rotates the source volume under every
proposal, runs `forward_3d2d`, correlates and takes `torch.max` at batch 1 with a `.item()` per pair
(test_co3d.py:133-152); here pairs are accumulated into batches, scored by ONE fused call, and the geodesic
error, Acc@15 and Acc@30 are computed on the device with one host read per category.

Datasets are duck-typed exactly as the reference uses them: iterating a CO3D-style dataset yields metadata
dicts with `n` and `model_id`, and `dataset.get_data(sequence_name=..., ids=...)` returns `image` [k,3,H,W] and
`R` [k,3,3] (data_loader_co3d.py); a LINEMOD-style loader yields dicts with `src_img, ref_img, src_mask,
ref_mask, src_R, ref_R`.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import so3


def get_permutations(num_frames: int) -> torch.Tensor:
    """All ordered pairs (i, j), i != j (test_co3d.py:46-52)."""
    return torch.tensor([(i, j) for i in range(num_frames) for j in range(num_frames) if i != j])


def geodesic_deg(R_a: torch.Tensor, R_b: torch.Tensor) -> torch.Tensor:
    """arccos((tr(Ra^T Rb) - 1) / 2) in degrees (test_co3d.py:149-150, modules/model.py:198-200)."""
    s = ((R_a.reshape(-1, 9) * R_b.reshape(-1, 9)).sum(-1).clamp(-1, 3) - 1) / 2
    return torch.arccos(s) * 180.0 / math.pi


@torch.no_grad()
def evaluate_category(cfg, model, dataset, num_frames: int = 2, use_pbar: bool = False, device="cuda",
                      batch_pairs: int = 32, proposals: torch.Tensor | None = None, return_details: bool = False):
    """test_co3d.py:93-154 for one category's `dataset`.  Returns the angular error of every evaluated pair
    (numpy, in the reference's order: sequences in dataset order, permutations in `get_permutations` order).
    `model` is the drop-in CO3D `Estimator` (`forward(img_src, img_tgt)` and `predict_rotation`)."""
    device = torch.device(device)
    permutations = get_permutations(num_frames)
    if proposals is None:   # ONCE per category, from the CPU generator (test_co3d.py:106)
        proposals = so3.random_rotations(cfg["DATA"]["NUM_ROTA"], device=device)
    errors, picked, pend_img1, pend_img2, pend_gt = [], [], [], [], []

    def flush():
        if not pend_img1:
            return
        img1, img2, gt = torch.cat(pend_img1), torch.cat(pend_img2), torch.cat(pend_gt)
        pend_img1.clear(); pend_img2.clear(); pend_gt.clear()
        vol_src, vol_tgt = model(img1, img2)
        _, idx, R_best, _ = model.predict_rotation(vol_src, vol_tgt, proposals)   # one fused call for the batch
        errors.append(geodesic_deg(R_best[:, 0], gt))
        picked.append(idx[:, 0])

    iterable = dataset
    if use_pbar:
        try:
            from tqdm.auto import tqdm
            iterable = tqdm(dataset)
        except ImportError:
            pass
    pending = 0
    for metadata in iterable:
        n, sequence_name = metadata["n"], metadata["model_id"]
        key_frames = np.random.choice(n, num_frames, replace=False)           # test_co3d.py:112
        batch = dataset.get_data(sequence_name=sequence_name, ids=key_frames)
        images_permuted = batch["image"][permutations].to(device)              # [P,2,3,H,W]
        rotations_permuted = batch["R"][permutations].to(device)
        pend_img1.append(images_permuted[:, 0])
        pend_img2.append(images_permuted[:, 1])
        pend_gt.append(torch.bmm(rotations_permuted[:, 0].transpose(1, 2), rotations_permuted[:, 1]))   # :121-124
        pending += len(permutations)
        if pending >= batch_pairs:
            flush()
            pending = 0
    flush()
    err = torch.cat(errors).cpu().numpy() if errors else np.zeros(0, dtype=np.float32)   # the one host read
    if return_details:
        return err, torch.cat(picked).cpu().numpy() if picked else np.zeros(0, dtype=np.int64), proposals
    return err


def evaluate_pairwise(cfg=None, model=None, split="train", num_frames=2, print_results=True, use_pbar=False,
                      categories=(), dataset="co3d", get_dataset=None, device="cuda", batch_pairs: int = 32):
    """test_co3d.py:157-198: per-category mean error, Acc@30, Acc@15 (percent) plus their means.
    `get_dataset(cfg=..., category=..., split=..., dataset=...)` builds one category's dataset (the reference's
    helper of the same name, test_co3d.py:55-90, needs the CO3D files)."""
    if get_dataset is None:
        raise ValueError("pass get_dataset(cfg=, category=, split=, dataset=): the CO3D files are not part of this repository")
    errors, errors_15, errors_30 = {}, {}, {}
    for category in categories:
        ds = get_dataset(cfg=cfg, category=category, split=split, dataset=dataset)
        ang = evaluate_category(cfg, model, ds, num_frames=num_frames, use_pbar=use_pbar, device=device, batch_pairs=batch_pairs)
        errors[category] = np.mean(ang)
        errors_15[category] = 100 * np.mean(ang < 15)
        errors_30[category] = 100 * np.mean(ang < 30)
        if print_results:
            print(category + " err: %.2f || acc_30: %.2f || acc_15: %.2f " % (errors[category], errors_30[category], errors_15[category]))
    errors["mean"] = np.mean(list(errors.values()))
    errors_15["mean"] = np.mean(list(errors_15.values()))
    errors_30["mean"] = np.mean(list(errors_30.values()))
    if print_results:
        print("avg_err: %.2f || avg_acc_30: %.2f || avg_acc_15: %.2f " % (errors["mean"], errors_30["mean"], errors_15["mean"]))
    return errors, errors_30, errors_15


@torch.no_grad()
def test_category(cfg, model, dataloader, device="cuda", return_details: bool = False):
    """test_linemod.py:20-84 for one object's `dataloader` (batches of pairs).  Per batch: skip if a mask is
    smaller than SIZE_THR (:39-41), draw a fresh codebook `random_rotations(model.num_rota)` (:43), lift both
    images (`model.forward(img_src, mask_src, img_tgt, mask_tgt)` - the signature modules/model.py:65 actually
    has; the script's 6-argument call is stale), score the codebook AND the ground-truth rotation (:53-60, one
    fused call each), select, measure.  Returns (mean error, Acc@30 %, Acc@15 %, pred_Rs [n,9]); with
    `return_details` also the per-pair errors and gt similarities."""
    device = torch.device(device)
    thr = cfg["DATA"]["SIZE_THR"]
    pred_Rs, pred_errs, gt_sims = [], [], []
    for data in dataloader:
        data = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in data.items()}
        mask_src, mask_tgt = data["src_mask"], data["ref_mask"]
        if torch.any(mask_src.flatten(1).sum(dim=-1) < thr) or torch.any(mask_tgt.flatten(1).sum(dim=-1) < thr):
            print("Skip bad case")
            continue
        codebook = so3.random_rotations(model.num_rota, device=device)
        img_feat_src, img_feat_tgt = model(data["src_img"], mask_src, data["ref_img"], mask_tgt)
        gt_src_2_tgt_R = torch.bmm(data["ref_R"], torch.inverse(data["src_R"]))          # modules/model.py:180-181
        _, _, R_best, _ = model.predict_rotation(img_feat_src, img_feat_tgt, codebook)
        gt_sims.append(model.score_rotations(img_feat_src, img_feat_tgt, gt_src_2_tgt_R[:, None].contiguous())[:, 0])
        pred_errs.append(geodesic_deg(R_best[:, 0], gt_src_2_tgt_R))
        pred_Rs.append(R_best[:, 0].reshape(-1, 9))
    if not pred_errs:
        raise RuntimeError("every batch was skipped (masks below SIZE_THR)")
    pred_err = torch.cat(pred_errs)
    acc_30 = 100 * (pred_err < 30).float().mean().item()
    acc_15 = 100 * (pred_err < 15).float().mean().item()
    out = (pred_err.mean().item(), acc_30, acc_15, torch.cat(pred_Rs).cpu().numpy())
    if return_details:
        return out + (pred_err.cpu().numpy(), torch.cat(gt_sims).cpu().numpy())
    return out


@torch.no_grad()
def test_objaverse(cfg, model, dataloader, device="cuda", out_dir: str | None = None):
    """test_objaverse.py:14-46 without Lightning: what `pl.Trainer().test(model, dataloader)` does for this model -
    `model.test_step(batch, i)` for every batch (modules/model.py:168-209: skip of small masks, a fresh hypothesis
    set per step, fused verification, geodesic error appended to `step_outputs`, ground-truth distance to `gt_dis`,
    predicted rotation to `pred_Rs`) - followed by the script's statistics.  Returns (mean error, Acc@30 %, Acc@15 %,
    pred_Rs [n,9]); with `out_dir` also writes `objaverse_pred_Rs.txt` and appends to `result.txt` like the script."""
    device = torch.device(device)
    model.step_outputs.clear()
    for i, batch in enumerate(dataloader):
        batch = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in batch.items()}
        model.test_step(batch, i)
    if not model.step_outputs:
        raise RuntimeError("every batch was skipped (masks below SIZE_THR)")
    pred_err = torch.cat(model.step_outputs)
    acc_30 = 100 * (pred_err < 30).float().mean().item()
    acc_15 = 100 * (pred_err < 15).float().mean().item()
    err = pred_err.mean().item()
    pred_Rs = np.concatenate([np.asarray(r).reshape(-1, 9) for r in model.pred_Rs])
    if out_dir is not None:
        import os

        os.makedirs(out_dir, exist_ok=True)
        np.savetxt(os.path.join(out_dir, "objaverse_pred_Rs.txt"), pred_Rs)
        with open(os.path.join(out_dir, "result.txt"), "a") as f:
            f.write("err: %.2f || avg_acc_30: %.2f || avg_acc_15: %.2f \n" % (err, acc_30, acc_15))
    return err, acc_30, acc_15, pred_Rs
