"""Checkpoint loading helpers (reference ``models/utils.py:6-21`` semantics)."""
import torch


def exists(x):
    return x is not None


def load_ckpt_state_dict(ckpt_path):
    """``.safetensors`` -> flat tensor dict; anything else -> ``torch.load(...)["state_dict"]``."""
    if str(ckpt_path).endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(ckpt_path)
    return torch.load(ckpt_path, map_location="cpu")["state_dict"]
