"""Pretransform adapter (interface parity with reference ``models/pretransforms.py:6-91``)."""
import torch
from torch import nn


class Pretransform(nn.Module):
    def __init__(self, enable_grad: bool, io_channels: int, is_discrete: bool):
        super().__init__()
        self.is_discrete = is_discrete
        self.io_channels = io_channels
        self.encoded_channels = None
        self.downsampling_ratio = None
        self.enable_grad = enable_grad

    def encode(self, x):
        raise NotImplementedError

    def decode(self, z):
        raise NotImplementedError


class AutoencoderPretransform(Pretransform):
    """``encode = encode_audio(x) / scale``, ``decode = decode_audio(z * scale)``.

    ``model_half`` is accepted for config compatibility; the native convolutions always use
    16-bit operands with fp32 accumulation and return fp32, so it changes nothing."""

    def __init__(self, model, scale=1.0, model_half=False, iterate_batch=False, chunked=False):
        is_discrete = model.bottleneck is not None and model.bottleneck.is_discrete
        super().__init__(enable_grad=False, io_channels=model.io_channels, is_discrete=is_discrete)
        self.model = model
        self.model.requires_grad_(False).eval()
        self.scale = scale
        self.downsampling_ratio = model.downsampling_ratio
        self.io_channels = model.io_channels
        self.sample_rate = model.sample_rate
        self.model_half = model_half
        self.iterate_batch = iterate_batch
        self.encoded_channels = model.latent_dim
        self.chunked = chunked
        self.num_quantizers = None
        self.codebook_size = None

    def encode(self, x, **kwargs):
        z = self.model.encode_audio(x.float(), chunked=self.chunked, iterate_batch=self.iterate_batch, **kwargs)
        return z.float() / self.scale

    def decode(self, z, **kwargs):
        z = z.float() * self.scale
        return self.model.decode_audio(z, chunked=self.chunked, iterate_batch=self.iterate_batch, **kwargs).float()

    def load_state_dict(self, state_dict, strict=True):
        self.model.load_state_dict(state_dict, strict=strict)
