from refiners_b200.foundationals.clip.image_encoder import CLIPImageEncoder, CLIPImageEncoderG, CLIPImageEncoderH
from refiners_b200.foundationals.clip.text_encoder import CLIPTextEncoder, CLIPTextEncoderG, CLIPTextEncoderH, CLIPTextEncoderL
from refiners_b200.foundationals.clip.tokenizer import CLIPTokenizer

__all__ = [
    "CLIPTextEncoder", "CLIPTextEncoderL", "CLIPTextEncoderH", "CLIPTextEncoderG", "CLIPTokenizer",
    "CLIPImageEncoder", "CLIPImageEncoderG", "CLIPImageEncoderH",
]
