"""Drop-in Estimator for the Objaverse / LINEMOD configuration.

Keeps the inference API of `/root/reference/modules/model.py`: `Estimator(cfg)`,
attributes (`feature_extractor`, `feature_aligner`, `num_rota`, `step_outputs`,
`gt_dis`, `pred_Rs`), `feature_extraction`, `forward(img_src, mask_src, img_tgt,
mask_tgt)` (:65-76), `validation_step` (:118-158), `test_step` (:168-209) with the
same batch keys.  The hypothesis-and-verification idiom inside the steps
(:131-146, :184-196) is one fused call into lib3dahv_b200.
"""
from __future__ import annotations

import torch

from modules._estimator_base import EstimatorBase, _ahv, geodesic_deg

torch.set_float32_matmul_precision("highest")   # modules/model.py:26


class Estimator(EstimatorBase):
    def __init__(self, cfg, feature_extractor=None):
        super().__init__(cfg, feature_extractor)
        self.gt_dis = []
        self.pred_Rs = []

    def forward(self, img_src, mask_src, img_tgt, mask_tgt):
        if self.cfg["DATA"]["BG"] is False:      # mask the input image (modules/model.py:67-69)
            img_src = img_src * mask_src
            img_tgt = img_tgt * mask_tgt
        feat_src = self.feature_extraction(img_src)
        feat_tgt = self.feature_extraction(img_tgt)
        return self.feature_aligner.forward_2d3d(feat_src, feat_tgt, random_mask=False, mask_ratio=0.0)

    def training_step(self, batch, batch_idx):
        """modules/model.py:78-116 (InfoNCE over per-pair hypothesis sets, invalid pairs masked out)."""
        mask_src, mask_tgt = batch["src_mask"], batch["ref_mask"]
        img_src, img_tgt = batch["src_img"], batch["ref_img"]
        if self.cfg["DATA"]["BG"] is False:
            img_src, img_tgt = img_src * mask_src, img_tgt * mask_tgt
        gt_R = self._gt_rotations(batch["src_R"], batch["ref_R"])
        vol_src, vol_tgt = self.feature_aligner.forward_2d3d(
            self.feature_extraction(img_src), self.feature_extraction(img_tgt),
            random_mask=self.cfg["TRAIN"]["MASK"], mask_ratio=self.cfg["TRAIN"]["MASK_RATIO"])
        self.Rs = self._sample_training_rotations(gt_R)
        thr = self.cfg["DATA"]["SIZE_THR"]
        valid = ((mask_src.flatten(1).sum(dim=-1) > thr) * (mask_tgt.flatten(1).sum(dim=-1) > thr)).float()
        if "dis_init" in batch.keys():
            valid = valid * (batch["dis_init"] < self.cfg["DATA"]["VIEW_THR"]).float()
        loss = self.infoNCE_loss(vol_src, vol_tgt, self.Rs, gt_R) * valid
        loss = loss.sum() / valid.sum().clamp(min=1e-8)
        self.log("train_loss", loss.item(), on_step=True, on_epoch=True, prog_bar=True, logger=True, sync_dist=True)
        return loss

    def configure_optimizers(self):
        return self._optimizers(backbone_lr_scale=0.1, step_size=20)      # modules/model.py:212-219

    @torch.no_grad()
    def predict(self, img_src, mask_src, img_tgt, mask_tgt, sampled_R=None, k: int = 1):
        """Batched inference entry (new): images -> best rotation(s) per pair, no host syncs."""
        vol_src, vol_tgt = self.forward(img_src, mask_src, img_tgt, mask_tgt)
        return self.predict_rotation(vol_src, vol_tgt, sampled_R, k)

    def _gt_rotations(self, R_src, R_tgt):
        with torch.no_grad():
            return torch.bmm(R_tgt, torch.inverse(R_src))

    def validation_step(self, batch, batch_idx):
        img_feat_src, img_feat_tgt = self.forward(batch["src_img"], batch["src_mask"], batch["ref_img"], batch["ref_mask"])
        gt_R = self._gt_rotations(batch["src_R"], batch["ref_R"])
        _, _, R_best, _ = self.predict_rotation(img_feat_src, img_feat_tgt)
        self.last_gt_sim = self.score_rotations(img_feat_src, img_feat_tgt, gt_R[:, None].contiguous())[:, 0]
        geo_dis = geodesic_deg(R_best[:, 0], gt_R)
        acc15, acc30 = (geo_dis <= 15).float().mean(), (geo_dis <= 30).float().mean()
        self.log("val_acc_15", acc15.item(), on_step=True, on_epoch=True, prog_bar=True, logger=True, sync_dist=True)
        self.log("val_acc_30", acc30.item(), on_step=True, on_epoch=True, prog_bar=True, logger=True, sync_dist=True)
        self.step_outputs.append(geo_dis)

    def on_validation_epoch_end(self):
        self.step_outputs.clear()

    def test_step(self, batch, batch_idx):
        mask_src, mask_tgt = batch["src_mask"], batch["ref_mask"]
        thr = self.cfg["DATA"]["SIZE_THR"]
        if torch.any(mask_src.flatten(1).sum(dim=-1) < thr) or torch.any(mask_tgt.flatten(1).sum(dim=-1) < thr):
            print("Skip bad case")
            return 0
        R_src, R_tgt = batch["src_R"], batch["ref_R"]
        img_feat_src, img_feat_tgt = self.forward(batch["src_img"], mask_src, batch["ref_img"], mask_tgt)
        gt_R = self._gt_rotations(R_src, R_tgt)
        _, _, R_best, _ = self.predict_rotation(img_feat_src, img_feat_tgt)
        pred_R = R_best[:, 0]
        geo_dis = geodesic_deg(pred_R, gt_R)
        gt_dis = geodesic_deg(R_src, R_tgt)
        self.step_outputs.append(geo_dis)
        self.gt_dis.append(gt_dis)
        self.pred_Rs.append(pred_R.cpu().detach().numpy().reshape(-1))
        self.log("test_error", geo_dis.mean().item(), on_step=True, prog_bar=True, logger=True, sync_dist=True)
