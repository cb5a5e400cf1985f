from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.model import StableDiffusion_XL
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.unet import SDXLUNet

__all__ = ["SDXLUNet", "StableDiffusion_XL"]
