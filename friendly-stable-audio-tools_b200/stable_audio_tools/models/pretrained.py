"""Hub loader (interface parity with reference ``models/pretrained.py:9-26``)."""
import json

from .factory import create_model_from_config
from .utils import load_ckpt_state_dict


def get_pretrained_model(name: str):
    from huggingface_hub import hf_hub_download
    with open(hf_hub_download(name, filename="model_config.json", repo_type="model")) as f:
        model_config = json.load(f)
    model = create_model_from_config(model_config)
    try:
        ckpt = hf_hub_download(name, filename="model.safetensors", repo_type="model")
    except Exception:
        ckpt = hf_hub_download(name, filename="model.ckpt", repo_type="model")
    model.load_state_dict(load_ckpt_state_dict(ckpt))
    return model, model_config
