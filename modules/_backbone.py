"""2D backbone used by the drop-in Estimator (stays in PyTorch; SURVEY.md §8 out of scope).

The reference builds `DPT_SwinV2_T_256(pretrained=True)` from its vendored MiDaS
tree on top of timm's `swinv2_tiny_window16_256` and consumes only the hooked
stage-4 activation `[B,768,8,8]` (modules/model.py:33,39-41).  timm, MiDaS
weights and network access are unavailable here, so the named architecture is
instantiated from torchvision with random weights; any module with the same
`[B,3,256,256] -> [B,768,8,8]` contract can be injected instead.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class SwinV2TinyStage4(nn.Module):
    """SwinV2-T, patch 4, window 16, 256x256 input: stage-4 tokens as [B,768,8,8]."""

    def __init__(self):
        super().__init__()
        from torchvision.models.swin_transformer import PatchMergingV2, SwinTransformer, SwinTransformerBlockV2

        self.model = SwinTransformer(patch_size=[4, 4], embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
                                     window_size=[16, 16], block=SwinTransformerBlockV2, downsample_layer=PatchMergingV2)
        self.model.head = nn.Identity()

    def forward(self, img: torch.Tensor) -> torch.Tensor:
        return self.model.features(img).permute(0, 3, 1, 2).contiguous()   # NHWC tokens, pre-norm


def build_backbone() -> nn.Module:
    return SwinV2TinyStage4()


# ---- loading a reference checkpoint ---------------------------------------------------------------------
# A reference Lightning checkpoint (test_co3d.py:218) holds, besides `feature_aligner.*` (names identical here),
# the MiDaS wrapper around timm's swinv2_tiny_window16_256: `feature_extractor.pretrained.model.<timm name>` plus
# the DPT decoder `feature_extractor.scratch.*`, which the model never runs (only the hooked stage-4 activation
# is consumed, modules/model.py:39-41).  The stand-in above is the same architecture under torchvision's names.
def _timm_to_torchvision(name: str):
    """timm (0.6.x) swin-v2 parameter name -> torchvision name, or None for entries with no counterpart."""
    import re

    if name.startswith("patch_embed.proj."):
        return "features.0.0." + name.rsplit(".", 1)[1]
    if name.startswith("patch_embed.norm."):
        return "features.0.2." + name.rsplit(".", 1)[1]
    m = re.match(r"layers\.(\d)\.downsample\.(reduction|norm)\.(\w+)$", name)
    if m:
        return f"features.{2 * int(m.group(1)) + 2}.{m.group(2)}.{m.group(3)}"
    m = re.match(r"layers\.(\d)\.blocks\.(\d+)\.(.+)$", name)
    if m:
        rest = m.group(3).replace("mlp.fc1.", "mlp.0.").replace("mlp.fc2.", "mlp.3.")
        return f"features.{2 * int(m.group(1)) + 1}.{m.group(2)}.{rest}"
    if name.startswith("norm."):
        return name
    return None          # head.*, attn_mask buffers


def load_reference_state_dict(estimator: nn.Module, state_dict: dict) -> dict:
    """Load a reference `Estimator` state dict (or Lightning checkpoint's `state_dict`) into the drop-in model:
    `feature_aligner.*` strictly by name; the backbone by the timm -> torchvision mapping above when the
    estimator uses the SwinV2-T stand-in (timm keeps q_bias / v_bias separately: torchvision's qkv.bias is their
    concatenation around a zero k-bias).  Returns {"loaded": n, "skipped": [names with no counterpart]}."""
    sd = state_dict.get("state_dict", state_dict)
    fa = {k[len("feature_aligner."):]: v for k, v in sd.items() if k.startswith("feature_aligner.")}
    estimator.feature_aligner.load_state_dict(fa, strict=True)
    loaded, skipped = len(fa), []
    prefix = "feature_extractor.pretrained.model."
    bb = getattr(estimator.feature_extractor, "model", None)
    if not isinstance(estimator.feature_extractor, SwinV2TinyStage4):
        skipped += [k for k in sd if k.startswith("feature_extractor.")]
        return {"loaded": loaded, "skipped": skipped}
    own = bb.state_dict()
    new, qv = {}, {}
    for k, v in sd.items():
        if not k.startswith("feature_extractor."):
            continue
        if not k.startswith(prefix):
            skipped.append(k)                      # feature_extractor.scratch.* (DPT decoder, never executed)
            continue
        name = k[len(prefix):]
        if name.endswith("attn.q_bias") or name.endswith("attn.v_bias"):
            qv[name] = v
            continue
        tv = _timm_to_torchvision(name)
        if tv is None or tv not in own or tuple(own[tv].shape) != tuple(v.shape):
            skipped.append(k)
            continue
        new[tv] = v
    for name, q in qv.items():
        if not name.endswith("q_bias"):
            continue
        tv = _timm_to_torchvision(name.replace("attn.q_bias", "attn.qkv.bias"))
        vb = qv.get(name.replace("q_bias", "v_bias"))
        if tv in own and vb is not None:
            new[tv] = torch.cat([q, torch.zeros_like(q), vb])
    missing = [k for k in own if k not in new and not k.endswith(("relative_position_index", "relative_coords_table"))]
    bb.load_state_dict(new, strict=False)
    return {"loaded": loaded + len(new), "skipped": skipped, "backbone_missing": missing}
