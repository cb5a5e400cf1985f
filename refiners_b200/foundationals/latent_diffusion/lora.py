"""``SDLoraManager``: named sets of LoRA layers on a Stable Diffusion model (UNet and, when present, text encoder).

Contract (method names and signatures, checkpoint-key conventions of the CivitAI-style state dicts, the ordering and
routing rules, the assertion texts) from /root/reference/src/refiners/foundationals/latent_diffusion/lora.py:10-330.
A "LoRA" here is a *set* of ``fluxion.adapters.lora.Lora`` layers sharing one name.

On CUDA nothing else is needed for speed: every ``LoraAdapter`` the manager creates around a ``Linear`` is evaluated as one
GEMM against the cached merged weight, re-merged when ``set_scale`` / ``update_scales`` change a scale
(refiners_b200.backend.merged_lora_weight); a captured CUDA graph notices scale changes through the value epoch.
The text encoder is optional in this package (it runs once per prompt, outside the per-step hot path): LoRA layers
addressed to it are skipped when the model has none.
"""

from __future__ import annotations

from typing import Any

from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200.fluxion.adapters.lora import Lora, LoraAdapter, auto_attach_loras
from refiners_b200.foundationals.latent_diffusion.model import LatentDiffusionModel

# checkpoint-key fragment -> the UNet layer class whose descendants those keys address; keys without such a fragment
# go to everything else (the transformer blocks)
_UNET_ROUTES = {"res": "ResidualBlock", "downsample": "Downsample", "upsample": "Upsample"}
# closely related checkpoint keys sometimes come in another order than the model calls them (q, k, v - in, out):
# rank of a key suffix inside its group; unknown suffixes go last
_SUFFIX_RANK = {"q": 1, "k": 2, "v": 3, "in": 3, "out": 4, "out0": 4, "out_0": 4}
_SUFFIXES = {pattern.format(stem): rank for stem, rank in _SUFFIX_RANK.items() for pattern in ("_{}", "_{}_lora")}


class SDLoraManager:
    def __init__(self, target: LatentDiffusionModel) -> None:
        self.target = target

    # -- the two adaptable parts ---------------------------------------------------------------------
    @property
    def unet(self) -> fl.Chain:
        part = self.target.unet
        assert isinstance(part, fl.Chain)
        return part

    @property
    def clip_text_encoder(self) -> fl.Chain:
        part = self.target.clip_text_encoder
        assert isinstance(part, fl.Chain)
        return part

    def _parts(self) -> list[fl.Chain]:
        encoder = self.target.clip_text_encoder
        return [self.unet] if encoder is None else [self.unet, encoder]

    # -- inspection ----------------------------------------------------------------------------------------
    @property
    def loras(self) -> list[Lora[Any]]:
        return [lora for part in self._parts() for lora in part.layers(Lora)]

    @property
    def lora_adapters(self) -> list[LoraAdapter]:
        return [adapter for part in self._parts() for adapter in part.layers(LoraAdapter)]

    @property
    def names(self) -> list[str]:
        return list({lora.name for lora in self.loras})

    @property
    def scales(self) -> dict[str, float]:
        return {name: self.get_scale(name) for name in self.names}

    def get_loras_by_name(self, name: str, /) -> list[Lora[Any]]:
        return [lora for lora in self.loras if lora.name == name]

    def get_scale(self, name: str, /) -> float:
        found = {lora.scale for lora in self.get_loras_by_name(name)}
        assert len(found) == 1, "lora scales are not all the same"
        return found.pop()

    def get_lora_weights(self, name: str) -> dict[str, Tensor]:
        """The factors of LoRA ``name`` keyed ``<parent path>.<n>.<target class>.{down,up}.weight`` (n counts the adapted
        children of one parent)."""
        out: dict[str, Tensor] = {}
        for part in self._parts():
            previous: fl.Chain | None = None
            count = 0
            for adapter, parent in part.walk(LoraAdapter):
                lora = adapter.loras.get(name)
                if lora is None:
                    continue
                count = count + 1 if parent is previous else 1
                previous = parent
                prefix = f"{parent.get_path()}.{count}.{type(adapter.target).__name__}"
                out[f"{prefix}.down.weight"] = lora.down.weight
                out[f"{prefix}.up.weight"] = lora.up.weight
        return out

    # -- scales ------------------------------------------------------------------------------------------------
    def set_scale(self, name: str, scale: float, /) -> None:
        self.update_scales({name: scale})

    def update_scales(self, scales: dict[str, float], /) -> None:
        known = self.names
        assert all(name in known for name in scales), f"Scales keys must be a subset of {known}"
        for lora in self.loras:
            if lora.name in scales:
                lora.scale = scales[lora.name]

    # -- adding ----------------------------------------------------------------------------------------------------
    def add_loras(
        self,
        name: str,
        /,
        tensors: dict[str, Tensor],
        scale: float = 1.0,
        unet_inclusions: list[str] | None = None,
        unet_exclusions: list[str] | None = None,
        unet_preprocess: dict[str, str] | None = None,
        text_encoder_inclusions: list[str] | None = None,
        text_encoder_exclusions: list[str] | None = None,
    ) -> None:
        """Load one LoRA from a CivitAI-style state dict (``..._q.down.weight`` / ``.up.weight`` pairs; keys mention
        ``unet`` or ``text``, bare keys are taken to address the UNet)."""
        assert name not in self.names, f"LoRA {name} already exists"
        placed = {key: tensor.to(device=self.target.device, dtype=self.target.dtype) for key, tensor in tensors.items()}
        layers = Lora.from_dict(name, state_dict=placed)
        layers = {key: layers[key] for key in sorted(layers, key=self.sort_keys)}
        if not any("unet" in key or "text" in key for key in layers):
            layers = {f"unet_{key}": lora for key, lora in layers.items()}
        self.add_loras_to_unet(layers, include=unet_inclusions, exclude=unet_exclusions, preprocess=unet_preprocess)
        self.add_loras_to_text_encoder(layers, include=text_encoder_inclusions, exclude=text_encoder_exclusions)
        self.set_scale(name, scale)

    def add_loras_to_text_encoder(
        self, loras: dict[str, Lora[Any]], /, include: list[str] | None = None, exclude: list[str] | None = None,
        debug_map: list[tuple[str, str]] | None = None,
    ) -> None:
        if self.target.clip_text_encoder is None:
            return
        mine = {key: lora for key, lora in loras.items() if "text" in key}
        auto_attach_loras(mine, self.clip_text_encoder, exclude=exclude, include=include, debug_map=debug_map)

    def add_loras_to_unet(
        self, loras: dict[str, Lora[Any]], /, include: list[str] | None = None, exclude: list[str] | None = None,
        preprocess: dict[str, str] | None = None, debug_map: list[tuple[str, str]] | None = None,
    ) -> None:
        """Keys naming a routed part (``res`` / ``downsample`` / ``upsample``) attach only below the layer class of that
        route; all remaining keys attach anywhere EXCEPT below those classes (and the always-excluded timestep encoder)."""
        mine = {key: lora for key, lora in loras.items() if "unet" in key}
        banned = ["TimestepEncoder"] if exclude is None else exclude
        routes = dict(_UNET_ROUTES if preprocess is None else preprocess)
        if include is not None:
            routes = {fragment: cls for fragment, cls in routes.items() if cls in include}
        routes = {fragment: cls for fragment, cls in routes.items() if cls not in banned}

        routed = {key for key in mine if any(fragment in key for fragment in routes)}
        for fragment, cls in routes.items():
            group = {key: mine[key] for key in mine if key in routed and fragment in key}
            auto_attach_loras(group, self.unet, include=[cls], exclude=banned, debug_map=debug_map)
        rest = {key: lora for key, lora in mine.items() if key not in routed}
        auto_attach_loras(rest, self.unet, exclude=[*banned, *routes.values()], include=include, debug_map=debug_map)

    # -- removing --------------------------------------------------------------------------------------------------
    def remove_loras(self, *names: str) -> None:
        for adapter in self.lora_adapters:
            for name in names:
                adapter.remove_lora(name)
            if not adapter.loras:
                adapter.eject()

    def remove_all(self) -> None:
        for adapter in self.lora_adapters:
            adapter.eject()

    # -- checkpoint key ordering ---------------------------------------------------------------------------------
    @staticmethod
    def _pad(input: str, /, padding_length: int = 2) -> str:
        """Zero-pad the purely numeric ``_``-separated fields: ``foo_3_bar`` -> ``foo_03_bar``."""
        return "_".join(field.zfill(padding_length) if field.isdigit() else field for field in input.split("_"))

    @staticmethod
    def sort_keys(key: str, /) -> tuple[str, int]:
        """Sort key that reorders checkpoint keys only *within* a group of siblings: (padded key without its known
        suffix, rank of that suffix)."""
        suffix, rank = next(((sfx, r) for sfx, r in _SUFFIXES.items() if key.endswith(sfx)), ("", 5))
        return SDLoraManager._pad(key.removesuffix(suffix)), rank
