"""Audio helpers (reference ``utils/audio_utils.py:22-27`` semantics)."""
import torch


def float_to_int16_audio(x: torch.Tensor, maximize: bool = False):
    peak = x.abs().max().item()
    div = peak if maximize else max(peak, 1.0)
    return x.div(div).mul(32767).to(torch.int16).cpu()
