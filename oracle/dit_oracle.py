"""CPU restatement of the reference DiffusionTransformer forward.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain functional torch on CPU
(fp32 by default, fp64 on request) over a flat state dict whose keys are the
reference's own (relative to ``DiffusionTransformer``; strip the
``model.model.`` prefix of a full checkpoint).  Each function cites the
reference lines it follows (paths relative to /root/reference/stable_audio_tools).

Covers the ``continuous_transformer`` backbone in both ``global_cond_type``
modes ("prepend", "adaLN"), batched classifier-free guidance and CFG rescale.
Pinned against the real reference modules by tests/test_oracle_vs_reference.py
(build container) and tests/golden/dit_*.npz (anywhere).
"""
import math

import torch
import torch.nn.functional as F


def _lin(x, w, b=None):
    return F.linear(x, w, b)


# Optional emulation of 16-bit tensor-core operands (tests only; None = the reference's plain fp32 arithmetic).
# With torch.float16 / torch.bfloat16 set, every contraction that the native path runs on tensor cores rounds its
# two operands to that type: the Linear layers of the blocks (input activations and weights), project_in / out,
# to_cond_embed, and the attention core (q, k, v and the probabilities P before P V); accumulation, LayerNorm,
# softmax, RoPE, biases, the residual stream and the timestep / global-embedding MLPs stay fp32.  The error of
# this variant against the fp32 oracle is the floor any implementation with 16-bit operands sits on - the
# reference's own GPU path (fp16 autocast, inference/sampling.py:210) included.
_OPERAND_DTYPE = None


class operand_rounding:
    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global _OPERAND_DTYPE
        self.prev, _OPERAND_DTYPE = _OPERAND_DTYPE, self.dtype

    def __exit__(self, *exc):
        global _OPERAND_DTYPE
        _OPERAND_DTYPE = self.prev


def _rnd(t):
    return t if (_OPERAND_DTYPE is None or t is None) else t.to(_OPERAND_DTYPE).to(t.dtype)


def _lin16(x, w, b=None):
    """A Linear layer that the native path runs as a 16-bit-operand GEMM."""
    return F.linear(_rnd(x), _rnd(w), b)


def fourier_features(t, weight):
    """models/blocks.py:95-97: f = 2*pi*t @ W^T ; cat[cos f, sin f]."""
    f = 2 * math.pi * t[:, None] @ weight.T
    return torch.cat([f.cos(), f.sin()], dim=-1)


def rotary_freqs(seq_len, inv_freq):
    """models/transformer.py:130-155: positions 0..N-1 (integers, exact) times
    inv_freq in fp32, duplicated to [N, 2*len(inv_freq)]."""
    t = torch.arange(seq_len, device=inv_freq.device).to(torch.float32)
    freqs = torch.einsum("i,j->ij", t, inv_freq.to(torch.float32))
    return torch.cat((freqs, freqs), dim=-1)


def rotate_half(x):
    """models/transformer.py:158-161: [a, b] -> [-b, a] on the two halves."""
    half = x.shape[-1] // 2
    return torch.cat((-x[..., half:], x[..., :half]), dim=-1)


def apply_rotary(t, freqs):
    """models/transformer.py:164-183: partial rotary on the first rot_dim dims,
    fp32 math, cast back to the input dtype."""
    out_dtype = t.dtype
    rot_dim, seq_len = freqs.shape[-1], t.shape[-2]
    dtype = torch.promote_types(torch.promote_types(t.dtype, freqs.dtype), torch.float32)
    freqs = freqs[-seq_len:, :].to(dtype)
    t = t.to(dtype)
    t_rot, t_pass = t[..., :rot_dim], t[..., rot_dim:]
    t_rot = t_rot * freqs.cos() + rotate_half(t_rot) * freqs.sin()
    return torch.cat((t_rot.to(out_dtype), t_pass.to(out_dtype)), dim=-1)


def layer_norm(x, gamma, beta=None):
    """models/transformer.py:188-206: F.layer_norm, eps 1e-5, beta is a zero buffer."""
    return F.layer_norm(x, x.shape[-1:], weight=gamma, bias=beta)


def attention_core(q, k, v):
    """models/transformer.py:510-536 (the einsum path the CPU reference takes):
    GQA by repeat_interleave (head h uses kv head h // (H/Hkv)), scale 1/sqrt(d),
    fp32 softmax, no mask, non-causal."""
    h, kv_h = q.shape[1], k.shape[1]
    if h != kv_h:
        rep = h // kv_h
        k = k.repeat_interleave(rep, dim=1)
        v = v.repeat_interleave(rep, dim=1)
    scale = 1.0 / (q.shape[-1] ** 0.5)
    dots = torch.einsum("bhid,bhjd->bhij", _rnd(q), _rnd(k)) * scale
    attn = F.softmax(dots, dim=-1, dtype=torch.float32).type(dots.dtype)
    return torch.einsum("bhij,bhjd->bhid", _rnd(attn), _rnd(v))


def _heads(x, h):
    b, n, _ = x.shape
    return x.view(b, n, h, -1).permute(0, 2, 1, 3)


def self_attention(x, sd, pfx, dim_heads, freqs, qk_norm=False):
    """models/transformer.py:407-554, fused to_qkv branch (:430-431), optional cosine-sim
    normalisation of q and k (:433-436), RoPE (:438-452)."""
    h = x.shape[-1] // dim_heads
    q, k, v = _lin16(x, sd[pfx + "to_qkv.weight"]).chunk(3, dim=-1)
    q, k, v = (_heads(t, h) for t in (q, k, v))
    if qk_norm:
        q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
    if freqs is not None:
        # the reference forces q, k to fp32 here (:444-446); apply_rotary promotes
        # to at least fp32 itself, and the fp64 mode of this oracle keeps fp64.
        q = apply_rotary(q, freqs)
        k = apply_rotary(k, freqs)
    o = attention_core(q, k, v)
    o = o.permute(0, 2, 1, 3).reshape(x.shape[0], x.shape[1], -1)
    return _lin16(o, sd[pfx + "to_out.weight"])


def cross_attention(x, ctx, sd, pfx, dim_heads, qk_norm=False):
    """models/transformer.py:420-427 (separate to_q / to_kv), kv_heads =
    dim_context // dim_heads (:306-312); no RoPE when a context is given (:438)."""
    h = x.shape[-1] // dim_heads
    q = _heads(_lin16(x, sd[pfx + "to_q.weight"]), h)
    k, v = _lin16(ctx, sd[pfx + "to_kv.weight"]).chunk(2, dim=-1)
    kv_h = k.shape[-1] // dim_heads
    k, v = _heads(k, kv_h), _heads(v, kv_h)
    if qk_norm:                                                                  # transformer.py:433-436
        q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
    o = attention_core(q, k, v)
    o = o.permute(0, 2, 1, 3).reshape(x.shape[0], x.shape[1], -1)
    return _lin16(o, sd[pfx + "to_out.weight"])


def feed_forward(x, sd, pfx):
    """models/transformer.py:222-235 (GLU: value = first half, gate = second
    half, SiLU on the gate) and :270 (output Linear with bias)."""
    u = _lin16(x, sd[pfx + "ff.0.proj.weight"], sd.get(pfx + "ff.0.proj.bias"))
    val, gate = u.chunk(2, dim=-1)
    m = val * F.silu(gate)
    return _lin16(m, sd[pfx + "ff.2.weight"], sd.get(pfx + "ff.2.bias"))


def transformer_block(x, ctx, global_cond, sd, pfx, dim_heads, freqs, qk_norm=False):
    """models/transformer.py:656-702."""
    ssg_key = pfx + "to_scale_shift_gate.1.weight"
    has_cross = (pfx + "cross_attn.to_q.weight") in sd and ctx is not None
    if ssg_key in sd and global_cond is not None:
        # adaLN branch, :665-689
        ssg = _lin(F.silu(global_cond), sd[ssg_key]).unsqueeze(1)
        scale_self, shift_self, gate_self, scale_ff, shift_ff, gate_ff = ssg.chunk(6, dim=-1)
        res = x
        a = layer_norm(x, sd[pfx + "pre_norm.gamma"], sd.get(pfx + "pre_norm.beta"))
        a = a * (1 + scale_self) + shift_self
        a = self_attention(a, sd, pfx + "self_attn.", dim_heads, freqs, qk_norm)
        x = a * torch.sigmoid(1 - gate_self) + res
        if has_cross:
            a = layer_norm(x, sd[pfx + "cross_attend_norm.gamma"], sd.get(pfx + "cross_attend_norm.beta"))
            x = x + cross_attention(a, ctx, sd, pfx + "cross_attn.", dim_heads, qk_norm)
        res = x
        a = layer_norm(x, sd[pfx + "ff_norm.gamma"], sd.get(pfx + "ff_norm.beta"))
        a = a * (1 + scale_ff) + shift_ff
        a = feed_forward(a, sd, pfx + "ff.")
        x = a * torch.sigmoid(1 - gate_ff) + res
    else:
        # plain branch, :691-700
        a = layer_norm(x, sd[pfx + "pre_norm.gamma"], sd.get(pfx + "pre_norm.beta"))
        x = x + self_attention(a, sd, pfx + "self_attn.", dim_heads, freqs, qk_norm)
        if has_cross:
            a = layer_norm(x, sd[pfx + "cross_attend_norm.gamma"], sd.get(pfx + "cross_attend_norm.beta"))
            x = x + cross_attention(a, ctx, sd, pfx + "cross_attn.", dim_heads, qk_norm)
        a = layer_norm(x, sd[pfx + "ff_norm.gamma"], sd.get(pfx + "ff_norm.beta"))
        x = x + feed_forward(a, sd, pfx + "ff.")
    return x


def continuous_transformer(x, prepend, ctx, global_cond, sd, depth, dim_heads, hidden_states=None, qk_norm=False):
    """models/transformer.py:764-809: project_in, cat prepend, rotary table for
    the full length (prepend token = position 0), blocks, project_out."""
    pfx = "transformer."
    x = _lin16(x, sd[pfx + "project_in.weight"])
    if prepend is not None:
        x = torch.cat((prepend, x), dim=-2)
    freqs = None
    if (pfx + "rotary_pos_emb.inv_freq") in sd:
        freqs = rotary_freqs(x.shape[1], sd[pfx + "rotary_pos_emb.inv_freq"])
    for i in range(depth):
        x = transformer_block(x, ctx, global_cond, sd, f"{pfx}layers.{i}.", dim_heads, freqs, qk_norm)
        if hidden_states is not None:
            hidden_states.append(x)
    return _lin16(x, sd[pfx + "project_out.weight"])


def _mlp(x, sd, name):
    b0, b2 = sd.get(name + ".0.bias"), sd.get(name + ".2.bias")
    return _lin(F.silu(_lin(x, sd[name + ".0.weight"], b0)), sd[name + ".2.weight"], b2)


def dit_inner_forward(sd, cfg, x, t, cross_attn_cond=None, global_embed=None, hidden_states=None,
                      input_concat_cond=None, prepend_cond=None):
    """models/dit.py:135-226 (``_forward``), continuous_transformer backbone, optional patching
    (dit.py:206-207,221-222), input_concat_cond (:163-168: nearest-neighbour resize to the latent length, channel
    concat before the 1x1 pre-conv) and prepend_cond (:157-161,185-195: to_prepend_embed MLP, its tokens in front of
    the global-conditioning token; "prepend" mode only)."""
    patch = cfg.get("patch_size", 1)
    depth = cfg["depth"]
    dim_heads = cfg["embed_dim"] // cfg["num_heads"]
    gtype = cfg.get("global_cond_type", "prepend")
    if cross_attn_cond is not None:
        cross_attn_cond = _lin16(F.silu(_lin16(cross_attn_cond, sd["to_cond_embed.0.weight"])),
                                 sd["to_cond_embed.2.weight"])                    # dit.py:149-150 (bias-free MLP)
    if global_embed is not None:
        global_embed = _mlp(global_embed, sd, "to_global_embed")                # dit.py:152-154
    te = _mlp(fourier_features(t, sd["timestep_features.weight"]), sd, "to_timestep_embed")  # dit.py:176
    global_embed = te if global_embed is None else global_embed + te             # dit.py:179-182
    prepend = None
    prepend_length = 0
    if prepend_cond is not None:
        prepend = _mlp(prepend_cond, sd, "to_prepend_embed")                     # dit.py:157-161 (bias-free MLP)
    if input_concat_cond is not None:                                            # dit.py:163-168
        if input_concat_cond.shape[2] != x.shape[2]:
            input_concat_cond = F.interpolate(input_concat_cond, (x.shape[2],), mode="nearest")
        x = torch.cat([x, input_concat_cond], dim=1)
    if gtype == "prepend":
        tok = global_embed.unsqueeze(1)                                          # dit.py:185-195
        prepend = tok if prepend is None else torch.cat([prepend, tok], dim=1)
        prepend_length = prepend.shape[1]
    x = F.conv1d(x, sd["preprocess_conv.weight"]) + x                            # dit.py:197
    x = x.transpose(1, 2)
    if patch > 1:                                                                # "b (t p) c -> b t (c p)"
        b, tp, c = x.shape
        x = x.reshape(b, tp // patch, patch, c).transpose(2, 3).reshape(b, tp // patch, c * patch)
    out = continuous_transformer(x, prepend, cross_attn_cond,
                                 global_embed if gtype == "adaLN" else None,
                                 sd, depth, dim_heads, hidden_states,
                                 qk_norm=bool(cfg.get("attn_kwargs", {}).get("qk_norm", False)))
    out = out.transpose(1, 2)[:, :, prepend_length:]                             # dit.py:219
    if patch > 1:                                                                # "b (c p) t -> b c (t p)"
        b, cp, tt = out.shape
        out = out.reshape(b, cp // patch, patch, tt).transpose(2, 3).reshape(b, cp // patch, tt * patch)
    return F.conv1d(out, sd["postprocess_conv.weight"]) + out                    # dit.py:224


def dit_forward(sd, cfg, x, t, cross_attn_cond=None, global_embed=None,
                negative_cross_attn_cond=None, negative_cross_attn_mask=None,
                cfg_scale=1.0, scale_phi=0.0, input_concat_cond=None, prepend_cond=None):
    """models/dit.py:228-364 (``forward``), eval mode: batched CFG (cond rows
    first, uncond rows second; null cond = zeros, or the negative cond),
    ``uncond + (cond - uncond) * cfg_scale`` and the optional std rescale."""
    if cfg_scale != 1.0 and (cross_attn_cond is not None or prepend_cond is not None):   # dit.py:270
        batch_cond = None
        if cross_attn_cond is not None:
            null = torch.zeros_like(cross_attn_cond)
            if negative_cross_attn_cond is not None:
                neg = negative_cross_attn_cond
                if negative_cross_attn_mask is not None:
                    neg = torch.where(negative_cross_attn_mask.to(torch.bool).unsqueeze(2), neg, null)
                batch_cond = torch.cat([cross_attn_cond, neg], dim=0)
            else:
                batch_cond = torch.cat([cross_attn_cond, null], dim=0)
        bx = torch.cat([x, x], dim=0)
        bt = torch.cat([t, t], dim=0)
        bg = None if global_embed is None else torch.cat([global_embed, global_embed], dim=0)
        bic = None if input_concat_cond is None else torch.cat([input_concat_cond, input_concat_cond], dim=0)   # :281-284
        bpc = None if prepend_cond is None else torch.cat([prepend_cond, torch.zeros_like(prepend_cond)], dim=0)  # :309-311
        out = dit_inner_forward(sd, cfg, bx, bt, batch_cond, bg, input_concat_cond=bic, prepend_cond=bpc)
        cond_out, uncond_out = torch.chunk(out, 2, dim=0)
        cfg_out = uncond_out + (cond_out - uncond_out) * cfg_scale
        if scale_phi != 0.0:
            cond_std = cond_out.std(dim=1, keepdim=True)
            cfg_std = cfg_out.std(dim=1, keepdim=True)
            return scale_phi * (cfg_out * (cond_std / cfg_std)) + (1 - scale_phi) * cfg_out
        return cfg_out
    return dit_inner_forward(sd, cfg, x, t, cross_attn_cond, global_embed, input_concat_cond=input_concat_cond,
                             prepend_cond=prepend_cond)


# ---------------------------------------------------------------------------
# synthetic weights (no checkpoints exist offline)
# ---------------------------------------------------------------------------

def dit_param_shapes(cfg):
    """Key -> shape for a DiffusionTransformer built from ``cfg`` (the
    ``model.diffusion.config`` dict of a reference JSON config), matching the
    reference state-dict layout."""
    D = cfg["embed_dim"]
    io = cfg["io_channels"]
    cin = io + cfg.get("input_concat_dim", 0)                                    # dit.py:38
    iop, cinp = io * cfg.get("patch_size", 1), cin * cfg.get("patch_size", 1)
    ct = cfg.get("cond_token_dim", 0)
    gd = cfg.get("global_cond_dim", 0)
    cond_embed = D if cfg.get("project_cond_tokens", True) else ct
    glob_embed = D if cfg.get("project_global_cond", True) else gd
    dh = D // cfg["num_heads"]
    shapes = {
        "timestep_features.weight": (128, 1),
        "to_timestep_embed.0.weight": (D, 256), "to_timestep_embed.0.bias": (D,),
        "to_timestep_embed.2.weight": (D, D), "to_timestep_embed.2.bias": (D,),
        "preprocess_conv.weight": (cin, cin, 1), "postprocess_conv.weight": (io, io, 1),
        "transformer.project_in.weight": (D, cinp), "transformer.project_out.weight": (iop, D),
        "transformer.rotary_pos_emb.inv_freq": (max(dh // 2, 32) // 2,),
    }
    if ct > 0:
        shapes["to_cond_embed.0.weight"] = (cond_embed, ct)
        shapes["to_cond_embed.2.weight"] = (cond_embed, cond_embed)
    if gd > 0:
        shapes["to_global_embed.0.weight"] = (glob_embed, gd)
        shapes["to_global_embed.2.weight"] = (glob_embed, glob_embed)
    if cfg.get("prepend_cond_dim", 0) > 0:
        shapes["to_prepend_embed.0.weight"] = (D, cfg["prepend_cond_dim"])
        shapes["to_prepend_embed.2.weight"] = (D, D)
    adaln = cfg.get("global_cond_type", "prepend") == "adaLN"
    for i in range(cfg["depth"]):
        p = f"transformer.layers.{i}."
        shapes[p + "pre_norm.gamma"] = (D,)
        shapes[p + "pre_norm.beta"] = (D,)
        shapes[p + "self_attn.to_qkv.weight"] = (3 * D, D)
        shapes[p + "self_attn.to_out.weight"] = (D, D)
        if ct > 0:
            shapes[p + "cross_attend_norm.gamma"] = (D,)
            shapes[p + "cross_attend_norm.beta"] = (D,)
            shapes[p + "cross_attn.to_q.weight"] = (D, D)
            shapes[p + "cross_attn.to_kv.weight"] = (2 * cond_embed, cond_embed)
            shapes[p + "cross_attn.to_out.weight"] = (D, D)
        shapes[p + "ff_norm.gamma"] = (D,)
        shapes[p + "ff_norm.beta"] = (D,)
        shapes[p + "ff.ff.0.proj.weight"] = (8 * D, D)
        shapes[p + "ff.ff.0.proj.bias"] = (8 * D,)
        shapes[p + "ff.ff.2.weight"] = (D, 4 * D)
        shapes[p + "ff.ff.2.bias"] = (D,)
        if adaln:
            shapes[p + "to_scale_shift_gate.1.weight"] = (6 * D, D)
    return shapes


def make_dit_weights(cfg, seed=0, std=0.02, dtype=torch.float32):
    """Deterministic synthetic weights.  Every tensor the reference zero-inits
    (to_out, ff.2, pre/postprocess_conv, to_scale_shift_gate) is
    re-randomised so parity is not vacuous; to_qkv / to_q / to_kv are scaled so
    the attention logits q.k/sqrt(d) have a standard deviation of about 2 for
    unit-variance inputs at any width (neither near-uniform nor one-hot softmax:
    std 0.02 would give ~0.6 at D=1536 and ~0.02 at D=256).  LN gamma ~ 1 + N(0, 0.1),
    beta = 0 (a buffer in the reference).  inv_freq follows transformer.py:115."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in dit_param_shapes(cfg).items():
        if k.endswith("inv_freq"):
            dim = 2 * shp[0]
            sd[k] = 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim))
        elif k.endswith(".gamma"):
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(".beta"):
            sd[k] = torch.zeros(shp)
        elif k == "timestep_features.weight":
            sd[k] = torch.randn(shp, generator=g) * 0.5            # FourierFeatures std 1, halved by DiTWrapper
        elif k.endswith("bias"):
            sd[k] = torch.randn(shp, generator=g) * std
        elif "to_qkv" in k or "to_q." in k or "to_kv" in k:
            # logits = q.k/8 with q, k ~ N(0, fan_in * s^2) per element: std = fan_in * s^2
            sd[k] = torch.randn(shp, generator=g) * math.sqrt(2.0 / shp[1])
        elif "preprocess_conv" in k or "postprocess_conv" in k:
            sd[k] = torch.randn(shp, generator=g) * (2.0 * std)
        else:
            sd[k] = torch.randn(shp, generator=g) * std
    return {k: v.to(dtype) if v.is_floating_point() else v for k, v in sd.items()}


def flops_per_row_per_block(N, D, M_ctx, ctx_dim, ff_inner):
    """SURVEY.md §8(d) algorithmic FLOPs of one TransformerBlock for one row."""
    self_f = 2 * N * D * 3 * D + 4 * N * N * D + 2 * N * D * D
    cross_f = 2 * N * D * D + 2 * M_ctx * ctx_dim * 2 * ctx_dim + 4 * N * M_ctx * D + 2 * N * D * D
    ff_f = 2 * N * D * 2 * ff_inner + 2 * N * ff_inner * D
    return self_f, cross_f, ff_f
