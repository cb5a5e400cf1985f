from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.controlnet import Controlnet, SD1ControlnetAdapter
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.model import StableDiffusion_1
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.self_attention_guidance import SD1SAGAdapter
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.unet import SD1UNet

__all__ = ["SD1UNet", "StableDiffusion_1", "Controlnet", "SD1ControlnetAdapter", "SD1SAGAdapter"]
