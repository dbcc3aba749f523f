"""CPU restatement of the reference Oobleck VAE (encoder, decoder, bottleneck).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional torch on CPU over a
flat state dict with the reference's keys relative to ``OobleckEncoder`` /
``OobleckDecoder`` (``layers.0.weight_g`` ...).

Third-party arithmetic: ``dac.nn.layers.WNConv1d/WNConvTranspose1d``
(descript-audio-codec 1.0.0, un-vendored) are ``torch.nn.utils.weight_norm``
wrappers, i.e. ``w = g * v / ||v||`` with the norm over all dims but 0 - torch's
own ``_weight_norm`` which *is* available; ``fold_weight_norm`` restates it.
"""
import math

import torch
import torch.nn.functional as F


def fold_weight_norm(weight_g, weight_v):
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v||_{dims != 0}.
    Conv1d weight [Cout, Cin, k] -> per-Cout norm; ConvTranspose1d weight
    [Cin, Cout, k] -> per-Cin norm (autoencoders.py:56,58,80,102 via dac)."""
    norm = weight_v.flatten(1).norm(dim=1).view(-1, *([1] * (weight_v.ndim - 1)))
    return weight_v * (weight_g / norm)


def snake_beta(x, alpha, beta):
    """models/blocks.py:318-319,350-358: log-scale alpha/beta per channel,
    x + sin^2(x * e^alpha) / (e^beta + 1e-9)."""
    a = torch.exp(alpha).view(1, -1, 1)
    b = torch.exp(beta).view(1, -1, 1)
    return x + (1.0 / (b + 0.000000001)) * torch.sin(x * a).pow(2)


# Optional emulation of the native path's 16-bit tensor-core operands (tests only): when set to
# torch.float16 / torch.bfloat16 every convolution rounds its input and its folded weight to that
# type (the accumulation, biases, Snake and the skip stream stay fp32, like the CUDA kernels).
# None (default) = the reference's plain fp32 arithmetic.
_OPERAND_DTYPE = None


class operand_rounding:
    """with operand_rounding(torch.float16): ... -> the oracle rounds conv operands like the GPU path."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global _OPERAND_DTYPE
        self.prev, _OPERAND_DTYPE = _OPERAND_DTYPE, self.dtype

    def __exit__(self, *exc):
        global _OPERAND_DTYPE
        _OPERAND_DTYPE = self.prev


def _rnd(t):
    return t if _OPERAND_DTYPE is None else t.to(_OPERAND_DTYPE).to(t.dtype)


def _wn_conv(x, sd, pfx, **kw):
    w = fold_weight_norm(sd[pfx + "weight_g"], sd[pfx + "weight_v"])
    return F.conv1d(_rnd(x), _rnd(w), sd.get(pfx + "bias"), **kw)


def _wn_convT(x, sd, pfx, **kw):
    w = fold_weight_norm(sd[pfx + "weight_g"], sd[pfx + "weight_v"])
    return F.conv_transpose1d(_rnd(x), _rnd(w), sd.get(pfx + "bias"), **kw)


def residual_unit(x, sd, pfx, dilation):
    """models/autoencoders.py:45-68: x + conv1(snake(conv7_dil(snake(x))))."""
    y = snake_beta(x, sd[pfx + "layers.0.alpha"], sd[pfx + "layers.0.beta"])
    y = _wn_conv(y, sd, pfx + "layers.1.", dilation=dilation, padding=(dilation * 6) // 2)
    y = snake_beta(y, sd[pfx + "layers.2.alpha"], sd[pfx + "layers.2.beta"])
    y = _wn_conv(y, sd, pfx + "layers.3.")
    return x + y


def oobleck_decoder(z, sd, cfg):
    """models/autoencoders.py:156-194 with DecoderBlock :88-116 (use_snake,
    transposed-conv upsampling, final_tanh optional)."""
    c_mults = [1] + list(cfg["c_mults"])
    strides = list(cfg["strides"])
    depth = len(c_mults)
    x = _wn_conv(z, sd, "layers.0.", padding=3)
    li = 1
    for i in range(depth - 1, 0, -1):
        s = strides[i - 1]
        p = f"layers.{li}."
        x = snake_beta(x, sd[p + "layers.0.alpha"], sd[p + "layers.0.beta"])
        x = _wn_convT(x, sd, p + "layers.1.", stride=s, padding=math.ceil(s / 2))
        for j, d in enumerate((1, 3, 9)):
            x = residual_unit(x, sd, f"{p}layers.{2 + j}.", d)
        li += 1
    x = snake_beta(x, sd[f"layers.{li}.alpha"], sd[f"layers.{li}.beta"])
    x = _wn_conv(x, sd, f"layers.{li + 1}.", padding=3)
    if cfg.get("final_tanh", True):
        x = torch.tanh(x)
    return x


def oobleck_encoder(a, sd, cfg):
    """models/autoencoders.py:119-153 with EncoderBlock :71-85."""
    c_mults = [1] + list(cfg["c_mults"])
    strides = list(cfg["strides"])
    depth = len(c_mults)
    x = _wn_conv(a, sd, "layers.0.", padding=3)
    li = 1
    for i in range(depth - 1):
        s = strides[i]
        p = f"layers.{li}."
        for j, d in enumerate((1, 3, 9)):
            x = residual_unit(x, sd, f"{p}layers.{j}.", d)
        x = snake_beta(x, sd[p + "layers.3.alpha"], sd[p + "layers.3.beta"])
        x = _wn_conv(x, sd, p + "layers.4.", stride=s, padding=math.ceil(s / 2))
        li += 1
    x = snake_beta(x, sd[f"layers.{li}.alpha"], sd[f"layers.{li}.beta"])
    return _wn_conv(x, sd, f"layers.{li + 1}.", padding=1)


def vae_sample(mean, scale, noise):
    """models/bottleneck.py:46-52 with the randn_like draw made explicit."""
    stdev = F.softplus(scale) + 1e-4
    return noise * stdev + mean


def vae_encode(h, noise):
    """models/bottleneck.py:59-62: mean, scale = chunk(2, dim=1)."""
    mean, scale = h.chunk(2, dim=1)
    return vae_sample(mean, scale, noise)


def reconstruct_audio_chunked(audio, esd, dsd, ecfg, dcfg, chunk_size, overlap, max_batch_size, noise_fn):
    """models/autoencoders.py:573-645 (AudioAutoencoder.reconstruct_audio, chunked=True) over the
    restated encoder / VAE bottleneck / decoder: zero-pad to n_chunk + 1 hops (:605-607), cut chunks of
    chunk_size latents hopping chunk_size - overlap (:609-616), encode -> vae_sample -> decode per batch of
    max_batch_size chunks (:619-625), Bartlett cross-fade on the shared edges (:631-640), crop (:643).
    noise_fn(mean) stands for the randn_like draw of bottleneck.py:50 (one call per encode batch)."""
    bs, n_ch, sample_length = audio.shape
    ratio = math.prod(ecfg["strides"])
    overlap_s = overlap * ratio
    win = torch.bartlett_window(overlap_s * 2)
    chunk_s = chunk_size * ratio
    hop = chunk_s - overlap_s
    n_chunk = int(math.ceil((sample_length - chunk_s) / hop)) + 1
    pad_len = chunk_s + hop * n_chunk - sample_length
    audio = F.pad(audio, (0, pad_len))
    chunks = torch.stack([audio[..., i * hop:i * hop + chunk_s] for i in range(n_chunk)], dim=1)
    chunks = chunks.reshape(bs * n_chunk, n_ch, chunk_s)
    xs = []
    for head in range(0, chunks.shape[0], max_batch_size):
        h = oobleck_encoder(chunks[head:head + max_batch_size], esd, ecfg)
        mean, scale = h.chunk(2, dim=1)
        z = vae_sample(mean, scale, noise_fn(mean))
        xs.append(oobleck_decoder(z, dsd, dcfg))
    xs = torch.cat(xs, dim=0)
    xs = xs.reshape(bs, n_chunk, xs.shape[1], xs.shape[2])
    rec = torch.zeros(bs, xs.shape[2], audio.shape[-1])
    for i in range(n_chunk):
        x_ = xs[:, i].clone()
        if i != 0:
            x_[:, :, :overlap_s] *= win[None, None, :overlap_s]
        if i != n_chunk - 1:
            x_[:, :, -overlap_s:] *= win[None, None, -overlap_s:]
        rec[:, :, i * hop:i * hop + chunk_s] += x_
    return rec[..., :sample_length]


# ---------------------------------------------------------------------------
# synthetic weights
# ---------------------------------------------------------------------------

def _conv_keys(shapes, pfx, cout, cin, k, bias=True, transposed=False):
    if transposed:
        shapes[pfx + "weight_g"] = (cin, 1, 1)
        shapes[pfx + "weight_v"] = (cin, cout, k)
    else:
        shapes[pfx + "weight_g"] = (cout, 1, 1)
        shapes[pfx + "weight_v"] = (cout, cin, k)
    if bias:
        shapes[pfx + "bias"] = (cout,)


def _res_keys(shapes, pfx, c):
    shapes[pfx + "layers.0.alpha"] = (c,)
    shapes[pfx + "layers.0.beta"] = (c,)
    _conv_keys(shapes, pfx + "layers.1.", c, c, 7)
    shapes[pfx + "layers.2.alpha"] = (c,)
    shapes[pfx + "layers.2.beta"] = (c,)
    _conv_keys(shapes, pfx + "layers.3.", c, c, 1)


def decoder_param_shapes(cfg):
    ch = cfg["channels"]
    c_mults = [1] + list(cfg["c_mults"])
    strides = list(cfg["strides"])
    shapes = {}
    _conv_keys(shapes, "layers.0.", c_mults[-1] * ch, cfg["latent_dim"], 7)
    li = 1
    for i in range(len(c_mults) - 1, 0, -1):
        cin, cout, s = c_mults[i] * ch, c_mults[i - 1] * ch, strides[i - 1]
        p = f"layers.{li}."
        shapes[p + "layers.0.alpha"] = (cin,)
        shapes[p + "layers.0.beta"] = (cin,)
        _conv_keys(shapes, p + "layers.1.", cout, cin, 2 * s, transposed=True)
        for j in range(3):
            _res_keys(shapes, f"{p}layers.{2 + j}.", cout)
        li += 1
    shapes[f"layers.{li}.alpha"] = (ch,)
    shapes[f"layers.{li}.beta"] = (ch,)
    _conv_keys(shapes, f"layers.{li + 1}.", cfg["out_channels"], ch, 7, bias=False)
    return shapes


def encoder_param_shapes(cfg):
    ch = cfg["channels"]
    c_mults = [1] + list(cfg["c_mults"])
    strides = list(cfg["strides"])
    shapes = {}
    _conv_keys(shapes, "layers.0.", ch, cfg["in_channels"], 7)
    li = 1
    for i in range(len(c_mults) - 1):
        cin, cout, s = c_mults[i] * ch, c_mults[i + 1] * ch, strides[i]
        p = f"layers.{li}."
        for j in range(3):
            _res_keys(shapes, f"{p}layers.{j}.", cin)
        shapes[p + "layers.3.alpha"] = (cin,)
        shapes[p + "layers.3.beta"] = (cin,)
        _conv_keys(shapes, p + "layers.4.", cout, cin, 2 * s)
        li += 1
    shapes[f"layers.{li}.alpha"] = (c_mults[-1] * ch,)
    shapes[f"layers.{li}.beta"] = (c_mults[-1] * ch,)
    _conv_keys(shapes, f"layers.{li + 1}.", cfg["latent_dim"], c_mults[-1] * ch, 3)
    return shapes


def decoder_transposed_prefixes(cfg):
    """Prefixes of the ConvTranspose1d layers of an OobleckDecoder state dict."""
    return {f"layers.{b}.layers.1." for b in range(1, len(cfg["c_mults"]) + 1)}


def make_oobleck_weights(shapes, seed=0, dtype=torch.float32, transposed=(), gain=0.7):
    """Deterministic synthetic weights.  weight_v ~ N(0,1); weight_g is set so
    the folded weight has element std gain/sqrt(fan_in) (keeps activations O(1)
    through ~35 layers) with a 10 % per-slice jitter so the fold is exercised;
    Snake alpha, beta ~ N(0, 0.3) (both are zero at init in the reference, which
    would make Snake parity vacuous: SURVEY.md H1); biases ~ N(0, 0.05).
    ``transposed`` = prefixes of ConvTranspose1d layers (weight_v [Cin,Cout,k],
    norm per Cin, effective fan-in Cin*k/stride = 2*Cin)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        if k.endswith("weight_v"):
            sd[k] = torch.randn(shp, generator=g)
        elif k.endswith("weight_g"):
            pfx = k[: -len("weight_g")]
            d0, d1, kk = shapes[pfx + "weight_v"]
            n_norm = d1 * kk
            fan_in = 2 * d0 if pfx in transposed else d1 * kk
            base = gain / math.sqrt(fan_in) * math.sqrt(n_norm)
            sd[k] = base * (1.0 + 0.1 * torch.randn(shp, generator=g))
        elif k.endswith(("alpha", "beta")):
            sd[k] = torch.randn(shp, generator=g) * 0.3
        else:
            sd[k] = torch.randn(shp, generator=g) * 0.05
    return {k: v.to(dtype) for k, v in sd.items()}
