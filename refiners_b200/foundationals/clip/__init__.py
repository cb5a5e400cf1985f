from refiners_b200.foundationals.clip.image_encoder import CLIPImageEncoder, CLIPImageEncoderG, CLIPImageEncoderH

__all__ = ["CLIPImageEncoder", "CLIPImageEncoderG", "CLIPImageEncoderH"]
