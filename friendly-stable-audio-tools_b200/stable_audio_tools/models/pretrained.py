"""Load a published model by its Hugging Face repo id (same entry point and return value as the
reference's ``models/pretrained.py``): the repo's JSON config is built with this package's
factory, then the checkpoint's state dict - reference key layout - is loaded into it."""
import json

from .factory import create_model_from_config
from .utils import load_ckpt_state_dict

_CONFIG_FILE = "model_config.json"
_CHECKPOINTS = ("model.safetensors", "model.ckpt")     # preferred first


def _fetch(repo_id, filename):
    from huggingface_hub import hf_hub_download      # imported lazily: offline users never need it
    return hf_hub_download(repo_id, filename=filename, repo_type="model")


def get_pretrained_model(name: str):
    """-> (model, model_config).  Needs network access or a warm Hugging Face cache."""
    with open(_fetch(name, _CONFIG_FILE)) as fh:
        config = json.load(fh)
    net = create_model_from_config(config)
    weights, last_err = None, None
    for candidate in _CHECKPOINTS:
        try:
            weights = _fetch(name, candidate)
            break
        except Exception as err:                     # the hub raises several unrelated types here
            last_err = err
    if weights is None:
        raise last_err
    net.load_state_dict(load_ckpt_state_dict(weights))
    return net, config
