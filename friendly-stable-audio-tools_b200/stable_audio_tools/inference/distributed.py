"""Data-parallel batched generation: the product-level form of the reference's ``generate.py:78-151``.

The path shards by prompt: every rank holds a full model replica and generates ``items[rank::world_size]``
(``generate.py:119-120``) in micro-batches; nothing on the data path is exchanged between ranks.  The single
collective is the one BASELINE.json's north star names: the conditioner (T5 + number embedders) runs ONCE, on rank 0,
for all prompts, and its output tensors are broadcast (NCCL over NVLink on GPUs; gloo in the CPU tests); every rank
then slices its shard out of the broadcast tensors.

Deviation from the reference, stated: the reference seeds once per batch (``generation.py:159-163``), so a prompt's
audio depends on which other prompts share its batch and on the world size.  ``per_prompt_seed=True`` (default)
derives the start noise and the per-step SDE noise of prompt i from ``seed + i`` (i = its index in the full
list), which makes the result of a prompt identical for every world size / batch size - the property the
1-vs-N test checks.  ``per_prompt_seed=False`` restores the reference's per-batch seeding.
"""
import typing as tp

import torch

from ..utils.torch_common import get_rank, get_world_size, shard_for_rank
from .generation import generate_diffusion_cond


class PerPromptNoise:
    """k-diffusion noise_sampler(sigma, sigma_next) -> unit normal [B, C, L] drawn from one generator per prompt."""

    def __init__(self, seeds: tp.Sequence[int], shape: tp.Sequence[int], device):
        self.shape, self.device = tuple(shape), device
        self.gens = [torch.Generator(device=device).manual_seed(int(s)) for s in seeds]

    def start_noise(self):
        return self()

    def __call__(self, sigma=None, sigma_next=None):
        return torch.stack([torch.randn(self.shape, generator=g, device=self.device) for g in self.gens], dim=0)


def broadcast_conditioning(cond: tp.Optional[dict], src: int = 0, device=None, template: tp.Optional[dict] = None):
    """Broadcast the conditioner output {id: (tensor, mask)} from rank `src` to every rank.  The structure (keys, shapes,
    dtypes) is sent first as one object broadcast, then one tensor broadcast per entry."""
    import torch.distributed as td
    if not (td.is_available() and td.is_initialized()) or td.get_world_size() == 1:
        return cond
    rank = td.get_rank()
    meta = [None]
    if rank == src:
        meta[0] = [(k, tuple(v[0].shape), str(v[0].dtype), tuple(v[1].shape), str(v[1].dtype)) for k, v in cond.items()]
    td.broadcast_object_list(meta, src=src)
    out = {}
    for key, shp, dt, mshp, mdt in meta[0]:
        if rank == src:
            t, m = cond[key][0].to(device).contiguous(), cond[key][1].to(device).contiguous()
        else:
            t = torch.empty(shp, dtype=getattr(torch, dt.split(".")[-1]), device=device)
            m = torch.empty(mshp, dtype=getattr(torch, mdt.split(".")[-1]), device=device)
        td.broadcast(t, src)
        if m.dtype == torch.bool:                     # NCCL has no bool broadcast
            mb = m.to(torch.uint8)
            td.broadcast(mb, src)
            m = mb.to(torch.bool)
        else:
            td.broadcast(m, src)
        out[key] = (t, m)
    return out


@torch.no_grad()
def generate_sharded(model, conditioning: tp.Optional[tp.List[dict]] = None, *, conditioning_tensors: tp.Optional[dict] = None,
                     steps: int = 100, cfg_scale: float = 7.0, sample_size: int = 2097152, batch_size: int = 4,
                     seed: int = 0, per_prompt_seed: bool = True, sampler_type: str = "dpmpp-3m-sde",
                     sigma_min: float = 0.3, sigma_max: float = 500.0, device="cuda", rank: tp.Optional[int] = None,
                     world_size: tp.Optional[int] = None, return_latents: bool = False, **kwargs):
    """Generate one clip per prompt, sharded ``rank::world_size`` over the process group.

    conditioning: the FULL list of per-prompt metadata dicts, identical on every rank (only rank 0 runs the
    conditioner), or conditioning_tensors: the conditioner output for the full list, valid on rank 0 (other ranks may
    pass None).  Returns [(global_index, tensor[channels, samples])] for this rank's prompts, in order.
    sigma_min / sigma_max default to the reference's generation values (generate.py:135-136)."""
    rank = get_rank() if rank is None else rank
    world_size = get_world_size() if world_size is None else world_size
    if conditioning_tensors is None:
        if conditioning is None:
            raise ValueError("need conditioning (metadata dicts) or conditioning_tensors")
        if rank == 0:
            model.conditioner.set_device(device)
            conditioning_tensors = model.conditioner(conditioning)
    cond_all = broadcast_conditioning(conditioning_tensors, 0, device)
    if cond_all is None:
        raise RuntimeError("no process group: every rank must be able to produce the conditioning itself")
    n_total = next(iter(cond_all.values()))[0].shape[0]
    mine = shard_for_rank(list(range(n_total)), rank, world_size)
    latent_len = sample_size // (model.pretransform.downsampling_ratio if model.pretransform else 1)
    results = []
    for b0 in range(0, len(mine), batch_size):
        idx = mine[b0:b0 + batch_size]
        sel = torch.tensor(idx, device=device)
        cond_b = {k: (v[0].to(device).index_select(0, sel), v[1].to(device).index_select(0, sel)) for k, v in cond_all.items()}
        extra = dict(kwargs)
        if per_prompt_seed:
            pn = PerPromptNoise([seed + i for i in idx], (model.io_channels, latent_len), device)
            extra.update(noise=pn.start_noise(), noise_sampler=pn)
            call_seed = seed
        else:
            call_seed = seed + b0 + rank          # one seed per batch call, like the reference
        out = generate_diffusion_cond(model, steps=steps, cfg_scale=cfg_scale, conditioning_tensors=cond_b,
                                      sample_size=sample_size, seed=call_seed, device=device, sampler_type=sampler_type,
                                      sigma_min=sigma_min, sigma_max=sigma_max, return_latents=return_latents,
                                      disable_tqdm=True, **extra)
        results.extend((i, out[k]) for k, i in enumerate(idx))
    return results
