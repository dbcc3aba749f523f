"""Bottlenecks on the hot path (reference ``models/bottleneck.py:10-65``).  The VAE
reparameterisation stays in PyTorch on purpose: it draws from the caller's torch RNG
(``randn_like``), and the kernel boundary is the deterministic ``mean | scale`` tensor
(SURVEY.md H8)."""
import torch
from torch import nn


class Bottleneck(nn.Module):
    def __init__(self, is_discrete: bool = False):
        super().__init__()
        self.is_discrete = is_discrete

    def encode(self, x, return_info=False, **kwargs):
        raise NotImplementedError

    def decode(self, x):
        raise NotImplementedError


class TanhBottleneck(Bottleneck):
    def encode(self, x, return_info=False):
        x = torch.tanh(x)
        return (x, {}) if return_info else x

    def decode(self, x):
        return x


def vae_sample(mean, scale):
    stdev = nn.functional.softplus(scale) + 1e-4
    var = stdev * stdev
    latents = torch.randn_like(mean) * stdev + mean
    kl = (mean * mean + var - torch.log(var) - 1).sum(1).mean()
    return latents, kl


class VAEBottleneck(Bottleneck):
    def encode(self, x, return_info=False, **kwargs):
        mean, scale = x.chunk(2, dim=1)
        z, kl = vae_sample(mean, scale)
        return (z, {"kl": kl}) if return_info else z

    def decode(self, x):
        return x
