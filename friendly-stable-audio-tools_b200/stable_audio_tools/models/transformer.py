"""Parameter containers for the continuous transformer of the DiT.

These classes reproduce the *interface* of reference ``models/transformer.py``
(class names, constructor kwargs, state-dict keys - SURVEY.md 3.3) so reference
checkpoints and JSON configs load unchanged.  They hold ``nn.Parameter``s only:
the arithmetic of ``ContinuousTransformer.forward`` (transformer.py:764-809) is
executed by ``libsatb200.so`` from ``DiffusionTransformer`` (models/dit.py here),
so calling ``forward`` on an inner container raises instead of silently running
eager PyTorch.
"""
import typing as tp

import torch
from torch import nn


class _FusedModule(nn.Module):
    """A module whose math is part of a fused native kernel sequence."""

    def forward(self, *args, **kwargs):
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container: its arithmetic is fused into the native "
            "DiffusionTransformer forward (libsatb200.so); call the enclosing DiffusionTransformer instead")


class RotaryEmbedding(_FusedModule):
    """Holds ``inv_freq`` (reference transformer.py:100-128). xpos / interpolation are not supported."""

    def __init__(self, dim, use_xpos=False, scale_base=512, interpolation_factor=1.0, base=10000,
                 base_rescale_factor=1.0):
        super().__init__()
        if use_xpos or interpolation_factor != 1.0:
            raise NotImplementedError("xpos / interpolated rotary embeddings are outside the native hot path")
        base = base * base_rescale_factor ** (dim / (dim - 2))
        self.register_buffer("inv_freq", 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim)))
        self.register_buffer("scale", None)
        self.dim = dim


class LayerNorm(_FusedModule):
    """gamma (parameter, or buffer when ``fix_scale``) and beta (buffer unless ``bias``)."""

    def __init__(self, dim, bias=False, fix_scale=False):
        super().__init__()
        if fix_scale:
            self.register_buffer("gamma", torch.ones(dim))
        else:
            self.gamma = nn.Parameter(torch.ones(dim))
        if bias:
            self.beta = nn.Parameter(torch.zeros(dim))
        else:
            self.register_buffer("beta", torch.zeros(dim))


class GLU(_FusedModule):
    """``proj``: Linear(dim_in, 2 * dim_out); value = first half, gate = second half (SiLU)."""

    def __init__(self, dim_in, dim_out, activation=None, use_conv=False, conv_kernel_size=3):
        super().__init__()
        if use_conv:
            raise NotImplementedError("convolutional GLU is outside the native hot path")
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(_FusedModule):
    """SwiGLU MLP; keys ``ff.0.proj.{weight,bias}`` and ``ff.2.{weight,bias}``."""

    def __init__(self, dim, dim_out=None, mult=4, no_bias=False, glu=True, use_conv=False, conv_kernel_size=3,
                 zero_init_output=True):
        super().__init__()
        if not glu or use_conv or no_bias or (dim_out not in (None, dim)) or mult != 4:
            raise NotImplementedError("only the default SwiGLU feed-forward (mult 4, biased) is on the native hot path")
        inner = int(dim * mult)
        linear_out = nn.Linear(inner, dim)
        if zero_init_output:
            nn.init.zeros_(linear_out.weight)
            nn.init.zeros_(linear_out.bias)
        self.ff = nn.Sequential(GLU(dim, inner), nn.Identity(), linear_out, nn.Identity())


class Attention(_FusedModule):
    """Fused ``to_qkv`` for self-attention, ``to_q`` + ``to_kv`` when a context dim is given."""

    def __init__(self, dim, dim_heads=64, dim_context=None, causal=False, zero_init_output=True, qk_norm=False,
                 natten_kernel_size=None):
        super().__init__()
        if causal or natten_kernel_size:
            raise NotImplementedError("causal / neighbourhood attention are outside the native hot path")
        self.dim, self.dim_heads = dim, dim_heads
        self.qk_norm = bool(qk_norm)      # cosine-similarity attention (reference transformer.py:433-436)
        dim_kv = dim_context if dim_context else dim
        self.num_heads = dim // dim_heads
        self.kv_heads = dim_kv // dim_heads
        if dim_context:
            self.to_q = nn.Linear(dim, dim, bias=False)
            self.to_kv = nn.Linear(dim_kv, dim_kv * 2, bias=False)
        else:
            self.to_qkv = nn.Linear(dim, dim * 3, bias=False)
        self.to_out = nn.Linear(dim, dim, bias=False)
        if zero_init_output:
            nn.init.zeros_(self.to_out.weight)


class TransformerBlock(_FusedModule):
    def __init__(self, dim, dim_heads=64, cross_attend=False, dim_context=None, global_cond_dim=None, causal=False,
                 zero_init_branch_outputs=True, conformer=False, layer_ix=-1, remove_norms=False, attn_kwargs={},
                 ff_kwargs={}, norm_kwargs={}):
        super().__init__()
        if conformer or remove_norms:
            raise NotImplementedError("conformer / norm-free blocks are outside the native hot path")
        self.dim, self.dim_heads = dim, dim_heads
        self.cross_attend, self.dim_context = cross_attend, dim_context
        self.global_cond_dim, self.layer_ix = global_cond_dim, layer_ix
        self.pre_norm = LayerNorm(dim, **norm_kwargs)
        self.self_attn = Attention(dim, dim_heads=dim_heads, causal=causal,
                                   zero_init_output=zero_init_branch_outputs, **attn_kwargs)
        if cross_attend:
            self.cross_attend_norm = LayerNorm(dim, **norm_kwargs)
            self.cross_attn = Attention(dim, dim_heads=dim_heads, dim_context=dim_context, causal=causal,
                                        zero_init_output=zero_init_branch_outputs, **attn_kwargs)
        self.ff_norm = LayerNorm(dim, **norm_kwargs)
        self.ff = FeedForward(dim, zero_init_output=zero_init_branch_outputs, **ff_kwargs)
        if global_cond_dim:
            self.to_scale_shift_gate = nn.Sequential(nn.SiLU(), nn.Linear(global_cond_dim, dim * 6, bias=False))
            nn.init.zeros_(self.to_scale_shift_gate[1].weight)


class ContinuousTransformer(_FusedModule):
    def __init__(self, dim, depth, *, dim_in=None, dim_out=None, dim_heads=64, cross_attend=False,
                 cond_token_dim=None, global_cond_dim=None, causal=False, rotary_pos_emb=True,
                 zero_init_branch_outputs=True, conformer=False, use_sinusoidal_emb=False, use_abs_pos_emb=False,
                 abs_pos_emb_max_length=10000, **kwargs):
        super().__init__()
        if causal or use_sinusoidal_emb or use_abs_pos_emb or not rotary_pos_emb:
            raise NotImplementedError("only the non-causal rotary configuration is on the native hot path")
        self.dim, self.depth = dim, depth
        self.project_in = nn.Linear(dim_in, dim, bias=False) if dim_in else nn.Identity()
        self.project_out = nn.Linear(dim, dim_out, bias=False) if dim_out else nn.Identity()
        self.rotary_pos_emb = RotaryEmbedding(max(dim_heads // 2, 32))
        self.layers = nn.ModuleList([
            TransformerBlock(dim, dim_heads=dim_heads, cross_attend=cross_attend, dim_context=cond_token_dim,
                             global_cond_dim=global_cond_dim, causal=causal,
                             zero_init_branch_outputs=zero_init_branch_outputs, conformer=conformer, layer_ix=i,
                             **kwargs)
            for i in range(depth)])
