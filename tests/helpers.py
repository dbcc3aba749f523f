"""Shared test helpers (CPU-side; the oracle is only ever the checker)."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_abs(a, b):
    return float((a.double() - b.double()).abs().max())


def build_native_dit(cfg, sd, device="cuda", operand_dtype="fp16"):
    """Drop-in DiffusionTransformer with the given flat state dict, on the GPU."""
    from stable_audio_tools.models.dit import DiffusionTransformer
    m = DiffusionTransformer(**cfg, operand_dtype=operand_dtype)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    return m.to(device).eval()


SAO_DIT = dict(io_channels=64, embed_dim=1536, depth=24, num_heads=24, cond_token_dim=768, global_cond_dim=1536,
               project_cond_tokens=False, transformer_type="continuous_transformer")
