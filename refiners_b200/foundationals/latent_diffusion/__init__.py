from refiners_b200.foundationals.latent_diffusion.auto_encoder import LatentDiffusionAutoencoder
from refiners_b200.foundationals.latent_diffusion.cross_attention import CrossAttentionBlock, CrossAttentionBlock2d
from refiners_b200.foundationals.latent_diffusion.model import LatentDiffusionModel
from refiners_b200.foundationals.latent_diffusion.range_adapter import RangeAdapter2d, RangeEncoder
from refiners_b200.foundationals.latent_diffusion.solvers import DDIM, Euler, Solver, SolverParams
from refiners_b200.foundationals.clip.text_encoder import CLIPTextEncoderL
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1 import (
    SD1Autoencoder,
    SD1ControlnetAdapter,
    SD1T2IAdapter,
    SD1UNet,
    StableDiffusion_1,
)
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl import (
    ControlLoraAdapter,
    DoubleTextEncoder,
    SDXLAutoencoder,
    SDXLT2IAdapter,
    SDXLUNet,
    StableDiffusion_XL,
)
from refiners_b200.foundationals.latent_diffusion.style_aligned import StyleAlignedAdapter
from refiners_b200.foundationals.latent_diffusion.unet_blocks import (
    ResidualAccumulator,
    ResidualBlock,
    ResidualConcatenator,
)

__all__ = [
    "CrossAttentionBlock", "CrossAttentionBlock2d", "LatentDiffusionModel", "RangeAdapter2d", "RangeEncoder",
    "DDIM", "Euler", "Solver", "SolverParams", "SD1UNet", "StableDiffusion_1", "SDXLUNet", "StableDiffusion_XL",
    "ResidualAccumulator", "ResidualBlock", "ResidualConcatenator", "LatentDiffusionAutoencoder",
    "SD1ControlnetAdapter", "ControlLoraAdapter", "SDXLIPAdapter", "SD1IPAdapter", "SD1T2IAdapter", "SDXLT2IAdapter",
    "SD1Autoencoder", "SDXLAutoencoder", "DoubleTextEncoder", "CLIPTextEncoderL", "StyleAlignedAdapter",
]


def __getattr__(name: str):
    if name in ("SDXLIPAdapter", "SD1IPAdapter"):
        from refiners_b200.foundationals.latent_diffusion import image_prompt

        return getattr(image_prompt, name)
    raise AttributeError(name)
