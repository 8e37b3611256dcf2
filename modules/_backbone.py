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
