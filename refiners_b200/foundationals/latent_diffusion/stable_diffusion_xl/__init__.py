from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.control_lora import ControlLora, ControlLoraAdapter
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.model import SDXLAutoencoder, StableDiffusion_XL
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.self_attention_guidance import SDXLSAGAdapter
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.t2i_adapter import SDXLT2IAdapter
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.text_encoder import DoubleTextEncoder
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.unet import SDXLUNet

__all__ = [
    "SDXLUNet", "StableDiffusion_XL", "SDXLAutoencoder", "DoubleTextEncoder", "ControlLora", "ControlLoraAdapter", "SDXLIPAdapter",
    "SDXLSAGAdapter", "SDXLT2IAdapter",
]


def __getattr__(name: str):  # SDXLIPAdapter lives next to IPAdapter (it pulls in the CLIP image tower): imported on demand
    if name == "SDXLIPAdapter":
        from refiners_b200.foundationals.latent_diffusion.image_prompt import SDXLIPAdapter

        return SDXLIPAdapter
    raise AttributeError(name)
