"""T2I-Adapter condition encoders, restated.  TEST INFRASTRUCTURE - see oracle/__init__.py.

Reference: /root/reference/src/refiners/foundationals/latent_diffusion/t2i_adapter.py
  Downsample2d :17-19 (2x2 average pooling), ResidualBlock :22-36 (x + conv1x1(relu(conv3x3(x)))), ResidualBlocks :39-63,
  ConditionEncoder :94-127 (pixel-unshuffle 8, stem, stages at 1/8, 1/16, 1/32, 1/64), ConditionEncoderXL :130-161
  (pixel-unshuffle 16, stages at 1/16, 1/16, 1/32, 1/32).  The features enter the UNets in oracle/unet.py (``t2i=``).
"""

from __future__ import annotations

from typing import Mapping

import torch.nn.functional as F
from torch import Tensor

from oracle import ops

SD = Mapping[str, Tensor]


def _conv(sd: SD, prefix: str, x: Tensor, padding: int = 0) -> Tensor:
    return ops.conv2d(x, sd[prefix + ".weight"], sd[prefix + ".bias"], padding=padding)


def condition_encoder(sd: SD, image: Tensor, xl: bool = False, num_residual_blocks: int = 2) -> tuple[Tensor, ...]:
    h = _conv(sd, "Conv2d", F.pixel_unshuffle(image, 16 if xl else 8), padding=1)
    halving = (False, False, True, False) if xl else (False, True, True, True)
    features = []
    for n, halve in enumerate(halving):
        p = f"StatefulResidualBlocks_{n + 1}.ResidualBlocks"
        if halve:
            h = F.avg_pool2d(h, 2)
        if (p + ".Conv2d.weight") in sd:
            h = _conv(sd, p + ".Conv2d", h)
        for k in range(num_residual_blocks):
            r = p + f".Chain.ResidualBlock_{k + 1}"
            h = h + _conv(sd, r + ".Conv2d_2", ops.relu(_conv(sd, r + ".Conv2d_1", h, padding=1)))
        features.append(h)
    return tuple(features)
