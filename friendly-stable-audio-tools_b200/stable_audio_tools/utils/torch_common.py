"""Rank / world-size helpers and the shape-tolerant state-dict copy
(reference ``utils/torch_common.py:12-25,46-61`` semantics)."""
import torch


def exists(x):
    return x is not None


def get_world_size():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_world_size()
    return 1


def get_rank():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank()
    return 0


def print_once(*args):
    if get_rank() == 0:
        print(*args)


def count_parameters(model):
    return sum(p.numel() for p in model.parameters()) + sum(b.numel() for b in model.buffers())


def copy_state_dict(model, state_dict):
    """Load only the entries whose key and shape match the model's."""
    own = model.state_dict()
    picked = {}
    for k, v in state_dict.items():
        if k in own and tuple(v.shape) == tuple(own[k].shape):
            picked[k] = v.data if isinstance(v, torch.nn.Parameter) else v
    own.update(picked)
    model.load_state_dict(own, strict=False)


def shard_for_rank(items, rank=None, world_size=None):
    """Round-robin data-parallel sharding of a work list, exactly the reference's
    ``items[rank::world_size]`` (``generate.py:119-120``, ``reconstruct_audios.py:118``)."""
    rank = get_rank() if rank is None else rank
    world_size = get_world_size() if world_size is None else world_size
    return items[rank::world_size]
