"""Conditional generation orchestrator (interface parity with reference
``inference/generation.py:94-261``): conditioning -> seeded noise -> k-diffusion sampler
over the native DiT -> native Oobleck decode."""
import math
import typing as tp

import numpy as np
import torch

from .sampling import sample_k, sample_rf
from .utils import prepare_audio


def build_mask(sample_size, mask_args):
    """Soft inpainting mask in [0, 1] (1 = keep the input audio), reference :268-290."""
    start = math.floor(mask_args["maskstart"] / 100.0 * sample_size)
    end = math.ceil(mask_args["maskend"] / 100.0 * sample_size)
    soft_l = round(mask_args["softnessL"] / 100.0 * sample_size)
    soft_r = round(mask_args["softnessR"] / 100.0 * sample_size)
    hann_l = torch.hann_window(soft_l * 2, periodic=False)[:soft_l]
    hann_r = torch.hann_window(soft_r * 2, periodic=False)[soft_r:]
    mask = torch.zeros((sample_size))
    mask[start:end] = 1
    mask[start:start + soft_l] = hann_l
    mask[end - soft_r:end] = hann_r
    if mask_args["marination"] > 0:
        mask = mask * (1 - mask_args["marination"])
    return mask


@torch.no_grad()
def generate_diffusion_cond(model, steps: int = 250, cfg_scale: float = 6,
                            conditioning: tp.Optional[tp.List[tp.Dict[str, tp.Any]]] = None,
                            conditioning_tensors: tp.Optional[dict] = None,
                            negative_conditioning: tp.Optional[tp.List[tp.Dict[str, tp.Any]]] = None,
                            negative_conditioning_tensors: tp.Optional[dict] = None, sample_size: int = 2097152,
                            seed: int = -1, device: str = "cuda",
                            init_audio: tp.Optional[tp.Tuple[int, torch.Tensor]] = None, init_noise_level: float = 1.0,
                            mask_args: dict = None, return_latents: bool = False, disable_tqdm: bool = False,
                            noise: tp.Optional[torch.Tensor] = None, **sampler_kwargs) -> torch.Tensor:
    """`noise` is the one argument the reference does not have: explicit unit-variance start noise
    [batch, io_channels, latent length] (inference/distributed.py seeds it per prompt so that a prompt's result does
    not depend on which rank / batch it lands in); None = the reference's seeded torch.randn draw."""
    if model.conditioner is not None:
        model.conditioner.set_device(device)
    audio_sample_size = sample_size
    if model.pretransform:
        sample_size //= model.pretransform.downsampling_ratio

    assert conditioning or conditioning_tensors, "Must provide either conditioning or conditioning_tensors"
    if conditioning_tensors is None:
        conditioning_tensors = model.conditioner(conditioning)
    conditioning_inputs = model.get_conditioning_inputs(conditioning_tensors)

    negative_inputs = {}
    if negative_conditioning or negative_conditioning_tensors:
        if negative_conditioning_tensors is None:
            negative_conditioning_tensors = model.conditioner(negative_conditioning)
        negative_inputs = model.get_conditioning_inputs(negative_conditioning_tensors, negative=True)

    num_sample = list(conditioning_tensors.values())[0][0].shape[0]   # batch size comes from the conditioning

    seed = seed if seed != -1 else np.random.randint(0, 2**32 - 1, dtype=np.uint32)
    torch.manual_seed(int(seed))
    if noise is None:
        noise = torch.randn([num_sample, model.io_channels, sample_size], device=device)
    elif tuple(noise.shape) != (num_sample, model.io_channels, sample_size):
        raise ValueError(f"noise must have shape {(num_sample, model.io_channels, sample_size)}, got {tuple(noise.shape)}")
    else:
        noise = noise.to(device)

    mask = None
    if init_audio is not None:
        in_sr, init_audio = init_audio
        io_channels = model.pretransform.io_channels if model.pretransform else model.io_channels
        init_audio = prepare_audio(init_audio, in_sr=in_sr, target_sr=model.sample_rate,
                                   target_length=audio_sample_size, target_channels=io_channels, device=device)
        if model.pretransform:
            init_audio = model.pretransform.encode(init_audio)
        init_audio = init_audio.repeat(num_sample, 1, 1)
        if mask_args is not None:
            crop_from = math.floor(mask_args["cropfrom"] / 100.0 * sample_size)
            paste_from = math.floor(mask_args["pastefrom"] / 100.0 * sample_size)
            paste_to = math.ceil(mask_args["pasteto"] / 100.0 * sample_size)
            assert paste_from < paste_to, "Paste From should be less than Paste To"
            crop_len = min(paste_to - paste_from, sample_size - crop_from)
            shifted = init_audio.new_zeros(init_audio.shape)
            shifted[:, :, paste_from:paste_from + crop_len] = init_audio[:, :, crop_from:crop_from + crop_len]
            init_audio = shifted
            mask = build_mask(sample_size, mask_args).to(device)
        else:
            sampler_kwargs["sigma_max"] = init_noise_level            # variation

    model_dtype = next(model.model.parameters()).dtype
    noise = noise.type(model_dtype)
    conditioning_inputs = {k: (v.type(model_dtype) if v is not None else v) for k, v in conditioning_inputs.items()}

    if model.diffusion_objective == "v":
        sampled = sample_k(model.model, noise, init_audio, mask, steps, **sampler_kwargs, **conditioning_inputs,
                           **negative_inputs, cfg_scale=cfg_scale, batch_cfg=True, rescale_cfg=True, device=device,
                           disable_tqdm=disable_tqdm)
    elif model.diffusion_objective == "rectified_flow":
        # discrete Euler on the velocity prediction (reference generation.py:236-246); no sigma_min / sampler choice
        rf_kwargs = {k: v for k, v in sampler_kwargs.items() if k not in ("sigma_min", "sampler_type")}
        sampled = sample_rf(model.model, noise, init_data=init_audio, steps=steps, **rf_kwargs, **conditioning_inputs,
                            **negative_inputs, cfg_scale=cfg_scale, batch_cfg=True, rescale_cfg=True, device=device,
                            disable_tqdm=disable_tqdm)
    else:
        raise NotImplementedError(f"diffusion objective '{model.diffusion_objective}'")
    del noise, conditioning_tensors, conditioning_inputs

    if model.pretransform and not return_latents:
        sampled = model.pretransform.decode(sampled.to(torch.float32))
    return sampled
