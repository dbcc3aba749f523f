"""Audio preparation helpers (reference ``inference/utils.py:7-39`` and ``PadCrop``,
``data/modification.py:12-24`` semantics)."""
import torch


def set_audio_channels(audio, target_channels):
    if target_channels == 1:
        return audio.mean(1, keepdim=True)
    if target_channels == 2:
        if audio.shape[1] == 1:
            return audio.repeat(1, 2, 1)
        if audio.shape[1] > 2:
            return audio[:, :2, :]
    return audio


def pad_crop(signal, n_samples):
    """Deterministic PadCrop: keep the first n_samples, zero-pad on the right."""
    n, s = signal.shape
    out = signal.new_zeros([n, n_samples])
    out[:, :min(s, n_samples)] = signal[:, :n_samples]
    return out


def prepare_audio(audio, in_sr, target_sr, target_length, target_channels, device):
    assert target_channels in (1, 2)
    audio = audio.to(device)
    if in_sr != target_sr:
        from torchaudio import transforms as T
        audio = T.Resample(in_sr, target_sr).to(device)(audio)
    audio = pad_crop(audio, target_length)
    if audio.dim() == 1:
        audio = audio[None, None]
    elif audio.dim() == 2:
        audio = audio[None]
    return set_audio_channels(audio, target_channels)
