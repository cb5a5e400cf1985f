"""SDLoraManager: named LoRA sets on a Stable Diffusion UNet.

Stand-alone behaviour (names, scales, removal, the "already exists" / "subset" assertions) and - where
/root/reference is mounted - the same operations on the real reference must yield the same module tree, the same
exported weight keys and the same checkpoint-key ordering."""

import sys
from pathlib import Path

import pytest
import torch

import refiners_b200.fluxion.layers as fl
from refiners_b200.foundationals.latent_diffusion import SD1UNet
from refiners_b200.foundationals.latent_diffusion.lora import SDLoraManager

BANNED = {"TimestepEncoder", "ResidualBlock", "Downsample", "Upsample"}


class Holder:
    """The three attributes of a LatentDiffusionModel the manager touches."""

    def __init__(self, unet, encoder=None):
        self.unet, self.clip_text_encoder = unet, encoder
        self.device, self.dtype = torch.device("meta"), torch.float32


def checkpoint(unet, layers_module) -> dict[str, torch.Tensor]:
    """A complete CivitAI-style LoRA state dict for the transformer-block Linears of ``unet`` (rank 4, meta tensors)."""
    tensors, i = {}, 0
    for lin, parent in unet.walk(layers_module.Linear):
        if {type(p).__name__ for p in [*parent.get_parents(), parent]} & BANNED:
            continue
        tensors[f"lora_unet_{i:04d}_x.down.weight"] = torch.empty(4, lin.in_features, device="meta")
        tensors[f"lora_unet_{i:04d}_x.up.weight"] = torch.empty(lin.out_features, 4, device="meta")
        i += 1
    return tensors


def test_manager_add_scale_remove():
    unet = SD1UNet(4, device="meta")
    manager = SDLoraManager(Holder(unet))
    pristine = repr(unet)
    tensors = checkpoint(unet, fl)
    assert manager.names == [] and manager.scales == {}
    manager.add_loras("style", tensors=tensors, scale=0.4)
    manager.add_loras("subject", tensors)
    assert set(manager.names) == {"style", "subject"} and manager.scales == {"style": 0.4, "subject": 1.0}
    assert len(manager.lora_adapters) == len(tensors) // 2 and len(manager.get_loras_by_name("style")) == len(tensors) // 2
    with pytest.raises(AssertionError, match="already exists"):
        manager.add_loras("style", tensors=tensors)
    with pytest.raises(AssertionError, match="subset"):
        manager.update_scales({"nobody": 1.0})
    manager.set_scale("style", 0.9)
    assert manager.get_scale("style") == 0.9
    exported = manager.get_lora_weights("style")
    assert len(exported) == len(tensors) and all(k.endswith((".down.weight", ".up.weight")) for k in exported)
    manager.remove_loras("style")
    assert manager.names == ["subject"]
    manager.remove_all()
    assert manager.names == [] and repr(unet) == pristine


@pytest.mark.skipif(not Path("/root/reference/src").exists(), reason="/root/reference is not mounted here")
def test_manager_matches_the_reference():
    from oracle.pin_against_reference import _import_reference

    rfl = _import_reference()
    from refiners.foundationals.latent_diffusion.lora import SDLoraManager as RefManager
    from refiners.foundationals.latent_diffusion.stable_diffusion_1.unet import SD1UNet as RefUNet

    for key in ("lora_unet_down_blocks_1_attentions_0_transformer_blocks_0_attn1_to_q", "lora_unet_down_blocks_10_attentions_0_proj_in",
                "a_to_out_0_lora", "plain_key", "lora_te_text_model_encoder_layers_3_self_attn_k_proj", "x_in", "y_out0"):
        assert SDLoraManager.sort_keys(key) == RefManager.sort_keys(key), key
    ours, theirs = SD1UNet(4, device="meta"), RefUNet(4, device="meta")
    mine, ref = SDLoraManager(Holder(ours)), RefManager(Holder(theirs, rfl.Chain()))
    tensors = checkpoint(theirs, rfl)
    assert list(tensors) == list(checkpoint(ours, fl))

    def tree(unet):  # the Lambda line prints a function signature whose annotations differ (jaxtyping): not structure
        return [line for line in repr(unet).splitlines() if "Lambda(compute_sinusoidal_embedding" not in line]

    for name, scale in (("a", 0.4), ("b", 1.0)):
        mine.add_loras(name, tensors=tensors, scale=scale)
        ref.add_loras(name, tensors=tensors, scale=scale)
    assert sorted(mine.names) == sorted(ref.names) and mine.scales == ref.scales
    assert list(mine.get_lora_weights("a")) == list(ref.get_lora_weights("a"))
    assert tree(ours) == tree(theirs)
    mine.remove_loras("a"); ref.remove_loras("a")
    assert tree(ours) == tree(theirs)
    mine.remove_all(); ref.remove_all()
    assert tree(ours) == tree(theirs)
