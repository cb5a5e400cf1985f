from refiners_b200.fluxion.context import ContextProvider
from refiners_b200.fluxion.utils import load_from_safetensors, manual_seed, no_grad, norm, pad, save_to_safetensors

__all__ = ["ContextProvider", "load_from_safetensors", "manual_seed", "no_grad", "norm", "pad", "save_to_safetensors"]
