"""Oobleck VAE: the reference's module interface over the native conv kernels.

Interface parity with reference ``models/autoencoders.py`` for the Oobleck path:
``ResidualUnit`` / ``EncoderBlock`` / ``DecoderBlock`` / ``OobleckEncoder`` /
``OobleckDecoder`` (:45-194, same constructor kwargs and state-dict keys incl. the
weight-norm ``weight_g`` / ``weight_v`` pairs), ``AudioAutoencoder`` (:234-645: encode /
decode with ``iterate_batch`` micro-batching, chunked ``encode_audio`` / ``decode_audio`` /
``reconstruct_audio`` with Bartlett cross-fades) and the config factories (:693-787).

The encoder / decoder ``forward`` run in ``libsatb200.so`` (``satb_oobleck_*``): tcgen05
implicit-GEMM convolutions with Snake fused into the producing epilogue.  The inner blocks
are parameter containers.  The chunking / cross-fade orchestration is host-side tensor
slicing on the device, exactly as in the reference.
"""
import ctypes
import math
import typing as tp

import torch
from torch import nn
from torch.nn import functional as F
from torch.nn.utils import weight_norm

from .. import _native
from .blocks import SnakeBeta
from .bottleneck import Bottleneck
from .factory import create_bottleneck_from_config, create_pretransform_from_config
from .transformer import _FusedModule


def WNConv1d(*args, **kwargs):
    """dac.nn.layers.WNConv1d: a weight-normed Conv1d (parameters weight_g, weight_v, bias)."""
    return weight_norm(nn.Conv1d(*args, **kwargs))


def WNConvTranspose1d(*args, **kwargs):
    return weight_norm(nn.ConvTranspose1d(*args, **kwargs))


def get_activation(activation: str, antialias=False, channels=None) -> nn.Module:
    if antialias:
        raise NotImplementedError("anti-aliased activations are outside the native hot path")
    if activation == "snake":
        return SnakeBeta(channels)
    if activation == "none":
        return nn.Identity()
    raise NotImplementedError(f"activation '{activation}' is outside the native hot path (use_snake=True only)")


class _Container(_FusedModule):
    pass


class ResidualUnit(_Container):
    def __init__(self, in_channels, out_channels, dilation, use_snake=False, antialias_activation=False):
        super().__init__()
        if not use_snake:
            raise NotImplementedError("only use_snake=True is on the native hot path")
        self.dilation = dilation
        self.layers = nn.Sequential(
            get_activation("snake", antialias=antialias_activation, channels=out_channels),
            WNConv1d(in_channels, out_channels, kernel_size=7, dilation=dilation, padding=(dilation * 6) // 2),
            get_activation("snake", antialias=antialias_activation, channels=out_channels),
            WNConv1d(out_channels, out_channels, kernel_size=1))


class EncoderBlock(_Container):
    def __init__(self, in_channels, out_channels, stride, use_snake=False, antialias_activation=False):
        super().__init__()
        self.layers = nn.Sequential(
            ResidualUnit(in_channels, in_channels, 1, use_snake=use_snake),
            ResidualUnit(in_channels, in_channels, 3, use_snake=use_snake),
            ResidualUnit(in_channels, in_channels, 9, use_snake=use_snake),
            get_activation("snake" if use_snake else "elu", antialias=antialias_activation, channels=in_channels),
            WNConv1d(in_channels, out_channels, kernel_size=2 * stride, stride=stride, padding=math.ceil(stride / 2)))


class DecoderBlock(_Container):
    def __init__(self, in_channels, out_channels, stride, use_snake=False, antialias_activation=False,
                 use_nearest_upsample=False):
        super().__init__()
        if use_nearest_upsample:
            raise NotImplementedError("nearest-neighbour upsampling is outside the native hot path")
        self.layers = nn.Sequential(
            get_activation("snake" if use_snake else "elu", antialias=antialias_activation, channels=in_channels),
            WNConvTranspose1d(in_channels, out_channels, kernel_size=2 * stride, stride=stride,
                              padding=math.ceil(stride / 2)),
            ResidualUnit(out_channels, out_channels, 1, use_snake=use_snake),
            ResidualUnit(out_channels, out_channels, 3, use_snake=use_snake),
            ResidualUnit(out_channels, out_channels, 9, use_snake=use_snake))


class _NativeOobleck(nn.Module):
    """Shared native-handle plumbing of OobleckEncoder / OobleckDecoder."""

    _is_decoder = False

    def _init_native(self, audio_channels, channels, latent_dim, c_mults, strides, final_tanh, operand_dtype):
        self.__dict__["_h"] = None
        self.__dict__["_dirty"] = True
        self.__dict__["_ncfg"] = dict(audio_channels=audio_channels, channels=channels, latent_dim=latent_dim,
                                      c_mults=list(c_mults), strides=list(strides), final_tanh=bool(final_tanh),
                                      operand_dtype=operand_dtype)
        # also fires when a parent module's load_state_dict recurses into this one (see models/dit.py)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.refresh_native_weights())

    def _apply(self, fn, *a, **k):
        self.__dict__["_dirty"] = True
        return super()._apply(fn, *a, **k)

    def refresh_native_weights(self):
        self.__dict__["_dirty"] = True

    def __del__(self):
        h = self.__dict__.get("_h")
        if h is not None:
            try:
                _native.lib().satb_oobleck_destroy(h)
            except Exception:
                pass

    def _handle(self, device):
        lib = _native.lib()
        nc = self.__dict__["_ncfg"]
        if self.__dict__["_h"] is None:
            cfg = _native.SatbOobleckConfig()
            cfg.in_channels, cfg.channels, cfg.latent_dim = nc["audio_channels"], nc["channels"], nc["latent_dim"]
            cfg.n_stages = len(nc["c_mults"])
            for i, (m, s) in enumerate(zip(nc["c_mults"], nc["strides"])):
                cfg.c_mults[i], cfg.strides[i] = m, s
            cfg.final_tanh = int(nc["final_tanh"])
            cfg.is_decoder = int(self._is_decoder)
            # "fp16" (default) | "bf16" | "fp16x3": split-operand mode, every conv product as (hi, hi) + (lo, hi) + (hi, lo)
            # with x_lo = fp16(x - x_hi): ~fp32 accuracy (the reference runs these convolutions in strict fp32,
            # inference/generation.py:165-166) at ~3x the tensor-core work
            if nc["operand_dtype"] not in ("fp16", "bf16", "fp16x3"):
                raise ValueError(f"operand_dtype must be fp16, bf16 or fp16x3, got {nc['operand_dtype']}")
            cfg.operand_dtype = {"fp16": 0, "bf16": 1, "fp16x3": 2}[nc["operand_dtype"]]
            h = ctypes.c_void_p()
            _native.check(lib.satb_oobleck_create(ctypes.byref(cfg), ctypes.byref(h)))
            self.__dict__["_h"] = h
        if self.__dict__["_dirty"]:
            st = _native.stream_ptr(device)
            with torch.no_grad():
                for name, t in self.state_dict().items():
                    if not t.is_cuda:
                        raise _native.NativeError(f"parameter {name} is on {t.device}: move the model to a CUDA device "
                                                  "(this package has no CPU path)")
                    src = t.detach().to(torch.float32).contiguous()
                    _native.check(lib.satb_oobleck_load_weight(self.__dict__["_h"], name.encode(), _native.ptr(src),
                                                               src.numel(), st))
                _native.check(lib.satb_oobleck_finalize(self.__dict__["_h"], st))
            self.__dict__["_dirty"] = False
        return self.__dict__["_h"]


class OobleckEncoder(_NativeOobleck):
    _is_decoder = False

    def __init__(self, in_channels=2, channels=128, latent_dim=32, c_mults=[1, 2, 4, 8], strides=[2, 4, 8, 8],
                 use_snake=False, antialias_activation=False, operand_dtype="fp16"):
        super().__init__()
        if not use_snake or antialias_activation:
            raise NotImplementedError("only use_snake=True without anti-aliasing is on the native hot path")
        self._init_native(in_channels, channels, latent_dim, c_mults, strides, False, operand_dtype)
        cm = [1] + list(c_mults)
        self.depth = len(cm)
        layers = [WNConv1d(in_channels, cm[0] * channels, kernel_size=7, padding=3)]
        for i in range(self.depth - 1):
            layers.append(EncoderBlock(cm[i] * channels, cm[i + 1] * channels, strides[i], use_snake=use_snake))
        layers += [get_activation("snake", channels=cm[-1] * channels),
                   WNConv1d(cm[-1] * channels, latent_dim, kernel_size=3, padding=1)]
        self.layers = nn.Sequential(*layers)
        self.downsampling_ratio = int(math.prod(strides))
        self.latent_dim = latent_dim

    @torch.no_grad()
    def forward(self, x):
        """audio [B, in_channels, T] -> pre-bottleneck [B, latent_dim, T / prod(strides)]"""
        if not x.is_cuda:
            raise _native.NativeError("OobleckEncoder.forward needs CUDA tensors (no CPU fallback)")
        with torch.cuda.device(x.device):            # the native handle / workspaces live on the model's device
            h = self._handle(x.device)
            xin = x.detach().to(torch.float32).contiguous()
            B, C, T = xin.shape
            out = torch.empty(B, self.latent_dim, T // self.downsampling_ratio, device=x.device, dtype=torch.float32)
            _native.check(_native.lib().satb_oobleck_encode(h, _native.ptr(xin), _native.ptr(out), B,
                                                            ctypes.c_longlong(T), _native.stream_ptr(x.device)))
        return out.to(x.dtype)


class OobleckDecoder(_NativeOobleck):
    _is_decoder = True

    def __init__(self, out_channels=2, channels=128, latent_dim=32, c_mults=[1, 2, 4, 8], strides=[2, 4, 8, 8],
                 use_snake=False, antialias_activation=False, use_nearest_upsample=False, final_tanh=True,
                 operand_dtype="fp16"):
        super().__init__()
        if not use_snake or antialias_activation:
            raise NotImplementedError("only use_snake=True without anti-aliasing is on the native hot path")
        self._init_native(out_channels, channels, latent_dim, c_mults, strides, final_tanh, operand_dtype)
        cm = [1] + list(c_mults)
        self.depth = len(cm)
        layers = [WNConv1d(latent_dim, cm[-1] * channels, kernel_size=7, padding=3)]
        for i in range(self.depth - 1, 0, -1):
            layers.append(DecoderBlock(cm[i] * channels, cm[i - 1] * channels, strides[i - 1], use_snake=use_snake,
                                       antialias_activation=antialias_activation,
                                       use_nearest_upsample=use_nearest_upsample))
        layers += [get_activation("snake", channels=cm[0] * channels),
                   WNConv1d(cm[0] * channels, out_channels, kernel_size=7, padding=3, bias=False),
                   nn.Tanh() if final_tanh else nn.Identity()]
        self.layers = nn.Sequential(*layers)
        self.upsampling_ratio = int(math.prod(strides))
        self.out_channels = out_channels

    @torch.no_grad()
    def forward(self, z):
        """latents [B, latent_dim, L] -> audio [B, out_channels, L * prod(strides)]"""
        if not z.is_cuda:
            raise _native.NativeError("OobleckDecoder.forward needs CUDA tensors (no CPU fallback)")
        with torch.cuda.device(z.device):            # the native handle / workspaces live on the model's device
            h = self._handle(z.device)
            zin = z.detach().to(torch.float32).contiguous()
            B, C, L = zin.shape
            out = torch.empty(B, self.out_channels, L * self.upsampling_ratio, device=z.device, dtype=torch.float32)
            _native.check(_native.lib().satb_oobleck_decode(h, _native.ptr(zin), _native.ptr(out), B, L,
                                                            _native.stream_ptr(z.device)))
        return out.to(z.dtype)


def _micro_batches(x, iterate_batch):
    """``iterate_batch`` (bool or int) is the micro-batch size: True -> 1 (reference :318-324)."""
    if not iterate_batch:
        return [x]
    bs = int(iterate_batch)
    return [x[i:i + bs] for i in range(0, x.shape[0], bs)]


class AudioAutoencoder(nn.Module):
    def __init__(self, encoder, decoder, latent_dim, downsampling_ratio, sample_rate, io_channels=2,
                 bottleneck: Bottleneck = None, pretransform=None, in_channels=None, out_channels=None,
                 soft_clip=False):
        super().__init__()
        self.downsampling_ratio = downsampling_ratio
        self.min_length = downsampling_ratio
        self.sample_rate = sample_rate
        self.latent_dim = latent_dim
        self.io_channels = io_channels
        self.in_channels = io_channels if in_channels is None else in_channels
        self.out_channels = io_channels if out_channels is None else out_channels
        self.encoder = encoder
        self.decoder = decoder
        self.bottleneck = bottleneck
        if pretransform is not None:
            raise NotImplementedError("nested pretransforms are outside the native hot path")
        self.pretransform = None
        self.soft_clip = soft_clip
        self.is_discrete = bool(self.bottleneck is not None and self.bottleneck.is_discrete)

    # -- plain encode / decode (reference :268-343) --------------------------------------
    def encode(self, audio, return_info=False, skip_pretransform=False, iterate_batch=False, **kwargs):
        latents = audio
        if self.encoder is not None:
            latents = torch.cat([self.encoder(a) for a in _micro_batches(audio, iterate_batch)], dim=0)
        info = {}
        if self.bottleneck is not None:
            latents, binfo = self.bottleneck.encode(latents, return_info=True, **kwargs)
            info.update(binfo)
        return (latents, info) if return_info else latents

    def decode(self, latents, iterate_batch=False, **kwargs):
        if self.bottleneck is not None:
            latents = torch.cat([self.bottleneck.decode(l) for l in _micro_batches(latents, iterate_batch)], dim=0)
        decoded = torch.cat([self.decoder(l) for l in _micro_batches(latents, iterate_batch)], dim=0)
        if self.soft_clip:
            decoded = torch.tanh(decoded)
        return decoded

    # -- chunked paths (reference :410-645) ----------------------------------------------
    @staticmethod
    def _chunk_starts(total, chunk, hop, extra=0):
        n_chunk = int(math.ceil((total - chunk) / hop)) + 1
        pad_len = chunk + hop * (n_chunk - 1 + extra) - total
        return n_chunk, pad_len

    @staticmethod
    def _crossfade_sum(pieces, n_chunk, hop, chunk, overlap, total_len, win):
        """Overlap-add chunk outputs [b, n_chunk, c, chunk] with a Bartlett fade on shared edges."""
        b, _, c, _ = pieces.shape
        out = torch.zeros((b, c, total_len), device=pieces.device)
        for i in range(n_chunk):
            piece = pieces[:, i]
            if i != 0:
                piece[:, :, :overlap] *= win[None, None, :overlap]
            if i != n_chunk - 1:
                piece[:, :, -overlap:] *= win[None, None, -overlap:]
            out[..., i * hop: i * hop + chunk] += piece
        return out

    def _run_chunks(self, chunks, fn, max_batch_size):
        outs = [fn(chunks[i:i + max_batch_size]) for i in range(0, chunks.shape[0], max_batch_size)]
        return torch.cat(outs, dim=0)

    def encode_audio(self, audio, chunked=False, chunk_size=128, overlap=4, max_batch_size=1, **kwargs):
        bs, n_ch, sample_length = audio.shape
        ratio = self.downsampling_ratio
        assert n_ch == self.in_channels
        assert sample_length % ratio == 0, "The audio length must be a multiple of compression ratio."
        if not chunked:
            return self.encode(audio, **kwargs)
        latent_length = sample_length // ratio
        hop_l = chunk_size - overlap
        win = torch.bartlett_window(overlap * 2, device=audio.device)
        chunk_s, hop_s = chunk_size * ratio, hop_l * ratio
        n_chunk, pad_len = self._chunk_starts(sample_length, chunk_s, hop_s)
        audio = F.pad(audio, (0, pad_len))                                   # zero padding
        chunks = torch.stack([audio[..., i * hop_s: i * hop_s + chunk_s] for i in range(n_chunk)], dim=1)
        zs = self._run_chunks(chunks.reshape(bs * n_chunk, n_ch, chunk_s), self.encode, max_batch_size)
        zs = zs.reshape(bs, n_chunk, zs.shape[1], zs.shape[2])
        latents = self._crossfade_sum(zs, n_chunk, hop_l, chunk_size, overlap, audio.shape[-1] // ratio, win)
        return latents[..., :latent_length]

    def decode_audio(self, latents, chunked=False, chunk_size=128, overlap=4, max_batch_size=1, **kwargs):
        bs, latent_dim, latent_length = latents.shape
        ratio = self.downsampling_ratio
        assert latent_dim == self.latent_dim
        if not chunked:
            return self.decode(latents, **kwargs)
        hop = chunk_size - overlap
        win = torch.bartlett_window(overlap * ratio * 2, device=latents.device)
        n_chunk, pad_len = self._chunk_starts(latent_length, chunk_size, hop)
        latents = F.pad(latents, (0, pad_len), mode="reflect")              # reflect padding
        chunks = torch.stack([latents[..., i * hop: i * hop + chunk_size] for i in range(n_chunk)], dim=1)
        xs = self._run_chunks(chunks.reshape(bs * n_chunk, latent_dim, chunk_size), self.decode, max_batch_size)
        xs = xs.reshape(bs, n_chunk, xs.shape[1], xs.shape[2])
        audio = self._crossfade_sum(xs, n_chunk, hop * ratio, chunk_size * ratio, overlap * ratio,
                                    latents.shape[-1] * ratio, win)
        return audio[..., :latent_length * ratio]

    @torch.no_grad()
    def reconstruct_audio(self, audio, chunked=True, chunk_size=128, overlap=4, max_batch_size=1, **kwargs):
        bs, n_ch, sample_length = audio.shape
        ratio = self.downsampling_ratio
        assert n_ch == self.in_channels
        if not chunked:
            return self.decode(self.encode(audio, **kwargs), **kwargs)
        win = torch.bartlett_window(overlap * ratio * 2, device=audio.device)
        chunk_s, overlap_s = chunk_size * ratio, overlap * ratio
        hop_s = chunk_s - overlap_s
        # the reference pads one hop more here than in encode_audio (:607 vs :455)
        n_chunk, pad_len = self._chunk_starts(sample_length, chunk_s, hop_s, extra=1)
        audio = F.pad(audio, (0, pad_len))
        chunks = torch.stack([audio[..., i * hop_s: i * hop_s + chunk_s] for i in range(n_chunk)], dim=1)
        xs = self._run_chunks(chunks.reshape(bs * n_chunk, n_ch, chunk_s), lambda c: self.decode(self.encode(c)),
                              max_batch_size)
        xs = xs.reshape(bs, n_chunk, xs.shape[1], xs.shape[2])
        rec = self._crossfade_sum(xs, n_chunk, hop_s, chunk_s, overlap_s, audio.shape[-1], win)
        return rec[..., :sample_length]


# ---------------------------------------------------------------------------------- factories
def create_encoder_from_config(encoder_config: tp.Dict[str, tp.Any]):
    if encoder_config["type"] != "oobleck":
        raise NotImplementedError(f"encoder '{encoder_config['type']}' is outside the native hot path (oobleck only)")
    encoder = OobleckEncoder(**encoder_config["config"])
    if not encoder_config.get("requires_grad", True):
        for p in encoder.parameters():
            p.requires_grad = False
    return encoder


def create_decoder_from_config(decoder_config: tp.Dict[str, tp.Any]):
    if decoder_config["type"] != "oobleck":
        raise NotImplementedError(f"decoder '{decoder_config['type']}' is outside the native hot path (oobleck only)")
    decoder = OobleckDecoder(**decoder_config["config"])
    if not decoder_config.get("requires_grad", True):
        for p in decoder.parameters():
            p.requires_grad = False
    return decoder


def create_autoencoder_from_config(config: tp.Dict[str, tp.Any]):
    ae = config["model"]
    bottleneck = ae.get("bottleneck")
    pretransform = ae.get("pretransform")
    if pretransform:
        pretransform = create_pretransform_from_config(pretransform, config["sample_rate"])
    return AudioAutoencoder(
        create_encoder_from_config(ae["encoder"]), create_decoder_from_config(ae["decoder"]),
        io_channels=ae["io_channels"], latent_dim=ae["latent_dim"], downsampling_ratio=ae["downsampling_ratio"],
        sample_rate=config["sample_rate"], bottleneck=create_bottleneck_from_config(bottleneck) if bottleneck else None,
        pretransform=pretransform, in_channels=ae.get("in_channels"), out_channels=ae.get("out_channels"),
        soft_clip=ae["decoder"].get("soft_clip", False))
