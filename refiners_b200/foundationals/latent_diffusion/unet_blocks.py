"""UNet building blocks shared by SD1.5 and SDXL.

Tree shapes follow /root/reference/src/refiners/foundationals/latent_diffusion/unet.py
(`ResidualBlock` :6-51, `ResidualAccumulator` :54-66, `ResidualConcatenator` :69-79).
"""

from __future__ import annotations

from typing import Any

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200 import backend as B
from refiners_b200.engine import fusion

Device = torch.device
DType = torch.dtype


class ResidualBlock(fl.Sum):
    """``(GN, SiLU, conv3x3, GN, SiLU, conv3x3)(x) + shortcut(x)``; the timestep bias is added
    later by injecting a RangeAdapter2d around the first conv."""

    def __init__(
        self, in_channels: int, out_channels: int, num_groups: int = 32, eps: float = 1e-5,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        if in_channels % num_groups != 0 or out_channels % num_groups != 0:
            raise ValueError("Number of input and output channels must be divisible by num_groups.")
        self.in_channels, self.out_channels, self.num_groups, self.eps = in_channels, out_channels, num_groups, eps
        kw = dict(device=device, dtype=dtype)
        # the shortcut is built first so that seeded random init draws in the reference's order
        shortcut = fl.Identity() if in_channels == out_channels else fl.Conv2d(in_channels, out_channels, kernel_size=1, **kw)
        body = fl.Chain(
            fl.GroupNorm(channels=in_channels, num_groups=num_groups, eps=eps, **kw),
            fl.SiLU(),
            fl.Conv2d(in_channels, out_channels, kernel_size=3, padding=1, **kw),
            fl.GroupNorm(channels=out_channels, num_groups=num_groups, eps=eps, **kw),
            fl.SiLU(),
            fl.Conv2d(out_channels, out_channels, kernel_size=3, padding=1, **kw),
        )
        super().__init__(body, shortcut)

    def forward(self, *inputs: Any) -> Any:
        """Sum(body, shortcut).  On CUDA the body's last conv adds shortcut(x) in its epilogue
        (one launch instead of conv + add); any other shape of the tree - e.g. an adapter
        spliced after the last conv - takes the generic Sum path."""
        if len(inputs) == 1 and isinstance(inputs[0], Tensor) and inputs[0].is_cuda and B.fusion_enabled() and len(self) == 2:
            body, shortcut = self[0], self[1]
            if type(body) is fl.Chain and body._steps()[-1][0] == "call" and not fusion._hooked(body, shortcut):
                last = fusion.tail_conv(body)
                if last is not None:
                    skip = shortcut(*inputs)
                    h = body._run_children(inputs, skip_last=True)
                    if isinstance(h, Tensor) and isinstance(skip, Tensor):
                        name = next(reversed(body._modules))
                        out = body._call_fused(name, lambda t: B.conv2d_module(t, last, residual=skip), h)
                        body._reset_context()
                        return out
                    raise RuntimeError("ResidualBlock: unexpected non-tensor intermediate")
        return fl.Sum.forward(self, *inputs)


class ResidualAccumulator(fl.Passthrough):
    """Add ``unet.residuals[n]`` to the activation and store the sum back at slot ``n``."""

    def __init__(self, n: int) -> None:
        self.n = n
        super().__init__(
            fl.Residual(fl.UseContext(context="unet", key="residuals").compose(lambda residuals: residuals[self.n])),
            fl.SetContext(context="unet", key="residuals", callback=self.update),
        )

    def update(self, residuals: list[Tensor | float], x: Tensor) -> None:
        residuals[self.n] = x


class ResidualConcatenator(fl.Chain):
    """Concatenate the skip connection ``unet.residuals[n]`` on the channel axis."""

    def __init__(self, n: int) -> None:
        self.n = n
        super().__init__(
            fl.Concatenate(
                fl.Identity(),
                fl.UseContext(context="unet", key="residuals").compose(lambda residuals: residuals[self.n]),
                dim=1,
            )
        )
