"""Generate tests/golden/resblock3d.npz by running the REFERENCE's own ResNetBlock_3D on CPU.

Runs only in the build container (needs /root/reference); the fixture is committed.

    python oracle/make_golden_lift.py

Imported unmodified from the reference (stub recipe of SURVEY.md Appendix A):
`modules.modules.ResNetBlock_3D` (modules/modules.py:9-47), configured as `Feature_Aligner` configures it
(:64: in_channels = mid_channel // 8 = 32, out_channels = 16, stride 1, BN=False) and applied to
`[m,32,8,8,8]` volumes as at :100-101.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.make_golden import import_reference  # noqa: E402


def main():
    import_reference()
    from modules.modules import ResNetBlock_3D      # the reference's (asserted inside import_reference)

    torch.manual_seed(0)
    block = ResNetBlock_3D(32, 16, stride=1, BN=False).eval()
    x = torch.randn(3, 32, 8, 8, 8) * 0.7 + 0.1      # the transformer's output scale
    with torch.no_grad():
        out = block(x.clone())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resblock3d.npz"), x=x.numpy(),
                        conv1_w=block.conv1.weight.detach().numpy(), conv2_w=block.conv2.weight.detach().numpy(),
                        down_w=block.downsample[0].weight.detach().numpy(), out=out.numpy())
    print("resblock3d.npz:", tuple(out.shape), float(out.abs().mean()))


if __name__ == "__main__":
    main()
