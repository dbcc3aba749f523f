"""FourierFeatures and SnakeBeta (reference ``models/blocks.py:88-97,330-358``)."""
import ctypes

import torch
from torch import nn

from .. import _native


class FourierFeatures(nn.Module):
    """Holds the random frequency matrix; evaluated inside the native DiT forward
    (``cat[cos f, sin f]``, ``f = 2*pi*t @ W^T``)."""

    def __init__(self, in_features, out_features, std=1.0):
        super().__init__()
        assert out_features % 2 == 0
        self.weight = nn.Parameter(torch.randn([out_features // 2, in_features]) * std)

    def forward(self, input):
        raise RuntimeError("FourierFeatures is fused into the native DiffusionTransformer forward")


class SnakeBeta(nn.Module):
    """``x + sin^2(alpha x) / (beta + 1e-9)`` with per-channel alpha, beta (log-scale by
    default); forward is the native ``satb_snake_beta`` kernel on [B, C, T] fp32."""

    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=True):
        super().__init__()
        self.in_features = in_features
        self.alpha_logscale = alpha_logscale
        init = torch.zeros(in_features) if alpha_logscale else torch.ones(in_features)
        self.alpha = nn.Parameter(init.clone() * alpha)
        self.beta = nn.Parameter(init.clone() * alpha)
        self.alpha.requires_grad = alpha_trainable
        self.beta.requires_grad = alpha_trainable
        self.no_div_by_zero = 0.000000001

    @torch.no_grad()
    def forward(self, x):
        if x.dim() != 3 or x.shape[1] != self.in_features:
            raise ValueError("SnakeBeta expects [B, C, T] with C == in_features")
        xc = x.float().contiguous()
        y = torch.empty_like(xc)
        B, C, T = xc.shape
        _native.check(_native.lib().satb_snake_beta(
            _native.dev_f32(xc, "x"), _native.dev_f32(self.alpha.detach().float().contiguous(), "alpha"),
            _native.dev_f32(self.beta.detach().float().contiguous(), "beta"), _native.dev_f32(y, "y"),
            B, C, ctypes.c_longlong(T), 1 if self.alpha_logscale else 0, _native.stream_ptr(x.device)))
        return y.to(x.dtype)
