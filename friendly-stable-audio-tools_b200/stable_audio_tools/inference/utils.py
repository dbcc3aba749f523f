"""Bring user audio into the shape the pipeline works on: [batch, channels, samples] at the model's
sample rate, exactly target_length long (behaviour of the reference's ``inference/utils.py``
``prepare_audio`` / ``set_audio_channels`` and of its deterministic ``PadCrop``)."""
import torch


def pad_crop(signal: torch.Tensor, n_samples: int) -> torch.Tensor:
    """[channels, samples] -> [channels, n_samples]: truncate on the right or zero-fill on the right."""
    kept = signal[:, :n_samples]
    if kept.shape[1] == n_samples:
        return kept.clone()
    return torch.nn.functional.pad(kept, (0, n_samples - kept.shape[1]))


def set_audio_channels(audio: torch.Tensor, target_channels: int) -> torch.Tensor:
    """[B, C, T] -> mono by averaging, stereo by duplicating a mono channel or dropping extra channels."""
    have = audio.shape[1]
    if target_channels == 1:
        return audio.mean(dim=1, keepdim=True)
    if target_channels == 2 and have == 1:
        return audio.expand(-1, 2, -1).contiguous()
    if target_channels == 2 and have > 2:
        return audio[:, :2]
    return audio


def prepare_audio(audio, in_sr, target_sr, target_length, target_channels, device):
    if target_channels not in (1, 2):
        raise AssertionError("target_channels must be 1 (mono) or 2 (stereo)")
    wav = audio.to(device)
    if in_sr != target_sr:
        from torchaudio.transforms import Resample
        wav = Resample(in_sr, target_sr).to(device)(wav)
    wav = pad_crop(wav, target_length)
    while wav.dim() < 3:                      # [T] -> [1, 1, T], [C, T] -> [1, C, T]
        wav = wav.unsqueeze(0)
    return set_audio_channels(wav, target_channels)
