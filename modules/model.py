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
