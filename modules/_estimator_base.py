"""Shared machinery of the drop-in Estimators (modules/model.py, modules/model_co3d.py)."""
from __future__ import annotations

import importlib
import math

import torch
import torch.nn as nn

try:  # Lightning is optional: the reference subclasses pl.LightningModule (modules/model.py:28)
    import lightning.pytorch as pl

    _Base = pl.LightningModule
except Exception:  # pragma: no cover - lightning is absent in this image
    pl = None
    _Base = nn.Module

from modules._backbone import build_backbone
from modules.modules import Feature_Aligner


def _ahv():
    return importlib.import_module("3dahv_b200")


def geodesic_deg(R_a: torch.Tensor, R_b: torch.Tensor) -> torch.Tensor:
    """arccos((tr(Ra^T Rb) - 1)/2) in degrees (modules/model.py:198-200)."""
    s = ((R_a.reshape(-1, 9) * R_b.reshape(-1, 9)).sum(-1).clamp(-1, 3) - 1) / 2
    return torch.arccos(s) * 180.0 / math.pi


class EstimatorBase(_Base):
    """Constructor, attributes and `feature_extraction` of the reference Estimator
    (modules/model.py:29-41); the hypothesis-and-verification idiom of the
    inference steps is replaced by one fused GPU call (`predict_rotation`)."""

    def __init__(self, cfg, feature_extractor: nn.Module | None = None):
        super().__init__()
        self.cfg = cfg
        self.num_rota = cfg["DATA"]["NUM_ROTA"]
        self.feature_extractor = build_backbone() if feature_extractor is None else feature_extractor
        self.feature_aligner = Feature_Aligner(in_channel=768, mid_channel=256, out_channel=32, n_heads=4, depth=4)
        self.step_outputs = []
        self._verifier = None

    if pl is None:
        def log(self, *args, **kwargs):  # Lightning's logger hook; a no-op without Lightning
            return None

    def feature_extraction(self, img):
        return self.feature_extractor(img)

    def verifier(self):
        """HypothesisVerifier bound to the CURRENT verification-head weights."""
        head = self.feature_aligner.feature_embedding_2d
        key = (head[0].weight.data_ptr(), head[0].weight._version, head[2].weight._version, head[2].bias._version,
               head[0].weight.device)
        if self._verifier is None or self._verifier[0] != key:
            self._verifier = (key, _ahv().HypothesisVerifier.from_feature_aligner(self.feature_aligner))
        return self._verifier[1]

    @torch.no_grad()
    def predict_rotation(self, img_feat_src, img_feat_tgt, sampled_R=None, k: int = 1):
        """modules/model.py:184-196 fused: sample (or take) the hypothesis set, score every
        hypothesis for every pair, select.  Returns (pred_sim_best [B,k], pred_index [B,k],
        pred_src_2_tgt_R [B,k,3,3], sampled_R)."""
        if sampled_R is None:
            sampled_R = _ahv().so3.random_rotations(self.num_rota, device=img_feat_src.device)
        r = self.verifier().score(img_feat_src, img_feat_tgt, sampled_R, k=k, return_scores=False)
        return r.topk_val, r.topk_idx, r.R_best, sampled_R

    @torch.no_grad()
    def score_rotations(self, img_feat_src, img_feat_tgt, R):
        """pred_sim for an explicit rotation set: R [N,3,3] shared or [B,N,3,3] per pair
        (the ground-truth hypothesis check of modules/model.py:137-143 is N=1 per pair)."""
        return self.verifier().score(img_feat_src, img_feat_tgt, R, k=1, return_scores=True).scores

    def infoNCE_loss(self, img_feat_1, img_feat_2, sampled_R, gt_delta_R):
        """Forward value of modules/model.py:43-63 (inference only — training and its
        backward pass are out of scope of this build)."""
        with torch.no_grad():
            acc = self.cfg["DATA"]["ACC_THR"]
            gt_sim = ((sampled_R.flatten(2) * gt_delta_R.reshape(-1, 1, 9)).sum(-1).clamp(-1, 3) - 1) / 2
            positive = (180 * torch.arccos(gt_sim) / math.pi) <= acc
            sim = self.score_rotations(img_feat_1, img_feat_2, sampled_R.contiguous())
            e = torch.exp(sim / 0.1)
            return -torch.log((e * positive).sum(-1) / e.sum(-1).clamp(min=1e-8))

    def training_step(self, batch, batch_idx):
        raise NotImplementedError("training is out of scope of the B200 hot-path build (SURVEY.md §8f-3)")
