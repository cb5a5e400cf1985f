from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.controlnet import Controlnet, SD1ControlnetAdapter
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.model import SD1Autoencoder, StableDiffusion_1
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.self_attention_guidance import SD1SAGAdapter
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.t2i_adapter import SD1T2IAdapter
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.unet import SD1UNet

__all__ = [
    "SD1UNet", "StableDiffusion_1", "SD1Autoencoder", "Controlnet", "SD1ControlnetAdapter", "SD1SAGAdapter", "SD1T2IAdapter",
    "SD1IPAdapter",
]


def __getattr__(name: str):  # the IP-Adapter lives next to IPAdapter (it pulls in the CLIP image tower): imported on demand
    if name == "SD1IPAdapter":
        from refiners_b200.foundationals.latent_diffusion.image_prompt import SD1IPAdapter

        return SD1IPAdapter
    raise AttributeError(name)
