from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.control_lora import ControlLoraAdapter
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.model import StableDiffusion_XL
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.self_attention_guidance import SDXLSAGAdapter
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.unet import SDXLUNet

__all__ = ["SDXLUNet", "StableDiffusion_XL", "ControlLoraAdapter", "SDXLIPAdapter", "SDXLSAGAdapter"]


def __getattr__(name: str):  # SDXLIPAdapter lives next to IPAdapter (it pulls in the CLIP image tower): imported on demand
    if name == "SDXLIPAdapter":
        from refiners_b200.foundationals.latent_diffusion.image_prompt import SDXLIPAdapter

        return SDXLIPAdapter
    raise AttributeError(name)
