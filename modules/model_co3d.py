"""Drop-in Estimator for the CO3D configuration.

Keeps the inference API of `/root/reference/modules/model_co3d.py`: `Estimator(cfg)`,
`feature_extraction`, `forward(img_src, img_tgt)` (:63-69).  The reference's CO3D
evaluation loop lives in `test_co3d.py:93-154`; its per-pair body (:133-152) is
`evaluate_pairs` here — batched, without per-pair host synchronisation.
"""
from __future__ import annotations

import torch

from modules._estimator_base import EstimatorBase, _ahv, geodesic_deg

torch.set_float32_matmul_precision("highest")   # modules/model_co3d.py:24


class Estimator(EstimatorBase):
    def __init__(self, cfg, feature_extractor=None):
        super().__init__(cfg, feature_extractor)
        self.mid_channel = 256

    def forward(self, img_src, img_tgt):
        feat_src = self.feature_extraction(img_src)
        feat_tgt = self.feature_extraction(img_tgt)
        return self.feature_aligner.forward_2d3d(feat_src, feat_tgt, random_mask=False, mask_ratio=0)

    def training_step(self, batch, batch_idx):
        """modules/model_co3d.py:71-96."""
        img_src, img_tgt = batch["image"][:, 0], batch["image"][:, 1]
        gt_R = batch["relative_rotation"].squeeze(1)
        vol_src, vol_tgt = self.feature_aligner.forward_2d3d(
            self.feature_extraction(img_src), self.feature_extraction(img_tgt),
            random_mask=self.cfg["TRAIN"]["MASK"], mask_ratio=self.cfg["TRAIN"]["MASK_RATIO"])
        sampled_R = self._sample_training_rotations(gt_R)
        loss = self.infoNCE_loss(vol_src, vol_tgt, sampled_R, gt_R).mean()
        self.log("train_loss", loss.item(), on_step=True, on_epoch=True, prog_bar=True, logger=True, sync_dist=True)
        return loss

    def configure_optimizers(self):
        return self._optimizers(backbone_lr_scale=1.0, step_size=200)     # modules/model_co3d.py:98-104

    @torch.no_grad()
    def predict(self, img_src, img_tgt, sampled_R=None, k: int = 1):
        vol_src, vol_tgt = self.forward(img_src, img_tgt)
        return self.predict_rotation(vol_src, vol_tgt, sampled_R, k)

    @torch.no_grad()
    def evaluate_pairs(self, img_src, img_tgt, gt_R, proposals=None):
        """test_co3d.py:133-152 for a batch of pairs: geodesic error (degrees) of the
        argmax hypothesis; `proposals` is the per-category rotation set of :106."""
        _, _, R_best, _ = self.predict(img_src, img_tgt, proposals)
        return geodesic_deg(R_best[:, 0], gt_R)
