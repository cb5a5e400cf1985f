"""Where the SAG probes go in an SD 1.5 UNet (contract:
/root/reference/src/refiners/foundationals/latent_diffusion/stable_diffusion_1/self_attention_guidance.py:12-41)."""

from __future__ import annotations

import refiners_b200.fluxion.layers as fl
from refiners_b200.fluxion.layers.attentions import ScaledDotProductAttention
from refiners_b200.foundationals.latent_diffusion.self_attention_guidance import SAGAdapter, SelfAttentionMap, SelfAttentionShape
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.unet import MiddleBlock, ResidualBlock, SD1UNet


def place_probes(middle_block: fl.Chain) -> None:
    """Shape probe after the first ResidualBlock; map probe in front of the first self-attention's SDPA."""
    middle_block.insert_after_type(ResidualBlock, SelfAttentionShape(context_key="middle_block_attn_shape"))
    attention = middle_block.ensure_find(fl.SelfAttention)
    attention.insert_before_type(
        ScaledDotProductAttention, SelfAttentionMap(num_heads=attention.num_heads, context_key="middle_block_attn_map")
    )


def remove_probes(middle_block: fl.Chain) -> None:
    middle_block.remove(middle_block.ensure_find(SelfAttentionShape))
    attention = middle_block.ensure_find(fl.SelfAttention)
    attention.remove(attention.ensure_find(SelfAttentionMap))


class SD1SAGAdapter(SAGAdapter[SD1UNet]):
    def __init__(self, target: SD1UNet, scale: float = 1.0, kernel_size: int = 9, sigma: float = 1.0) -> None:
        super().__init__(target=target, scale=scale, kernel_size=kernel_size, sigma=sigma)

    def inject(self: "SD1SAGAdapter", parent: fl.Chain | None = None) -> "SD1SAGAdapter":
        place_probes(self.target.ensure_find(MiddleBlock))
        return super().inject(parent)

    def eject(self) -> None:
        remove_probes(self.target.ensure_find(MiddleBlock))
        super().eject()
