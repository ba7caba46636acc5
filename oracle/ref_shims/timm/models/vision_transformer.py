"""checkpoint_filter_fn / _init_vit_weights restated (MolNexTR/models/transformers.py:16).
Only the initialiser runs on the reference path (weights are then overwritten by the seeded
synthetic checkpoint, so its exact distribution is irrelevant to parity)."""
import torch.nn as nn

from .layers import trunc_normal_


def checkpoint_filter_fn(state_dict, model):
    return state_dict


def _init_vit_weights(module, name="", head_bias=0.0, jax_impl=False):
    if isinstance(module, nn.Linear):
        trunc_normal_(module.weight, std=0.02)
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    elif isinstance(module, nn.LayerNorm):
        nn.init.zeros_(module.bias)
        nn.init.ones_(module.weight)
