from refiners_b200.foundationals.segment_anything.image_encoder import SAMViT, SAMViTH

__all__ = ["SAMViT", "SAMViTH"]
