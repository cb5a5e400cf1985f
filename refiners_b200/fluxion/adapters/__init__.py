from refiners_b200.fluxion.adapters.adapter import Adapter, lookup_top_adapter
from refiners_b200.fluxion.adapters.lora import Conv2dLora, LinearLora, Lora, LoraAdapter, auto_attach_loras

__all__ = ["Adapter", "lookup_top_adapter", "Lora", "LinearLora", "Conv2dLora", "LoraAdapter", "auto_attach_loras"]
