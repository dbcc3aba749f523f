"""Conditioners used by the Stable Audio text-to-audio configs (SURVEY.md 8f "next" row).

They run once per generation, not per denoise step, so they stay plain PyTorch modules:
``NumberConditioner`` (learned Fourier features of a normalised scalar, reference
``models/conditioners.py:64-102`` + ``models/adp.py:680-701,1495-1512``), ``IntConditioner``
(:39-61), ``T5Conditioner`` (frozen fp16 HF T5 encoder, padded to ``max_length``, masked
positions zeroed, :261-346) and ``MultiConditioner`` (:505-549); state-dict keys match the
reference (``conditioner.conditioners.<id>.embedder.embedding.0.weights`` ...).  The T5 encoder
needs the HF model files locally or a network, exactly like the reference.
"""
import logging
import math
import typing as tp
import warnings

import torch
from torch import nn


class Conditioner(nn.Module):
    def __init__(self, dim: int, output_dim: int, project_out: bool = False):
        super().__init__()
        self.dim, self.output_dim = dim, output_dim
        self.proj_out = nn.Linear(dim, output_dim) if (dim != output_dim or project_out) else nn.Identity()

    def set_device(self, device) -> None:
        raise NotImplementedError()


class LearnedPositionalEmbedding(nn.Module):
    """[x, sin(2 pi x w), cos(2 pi x w)] with learned frequencies w (continuous inputs)."""

    def __init__(self, dim: int):
        super().__init__()
        assert dim % 2 == 0
        self.weights = nn.Parameter(torch.randn(dim // 2))

    def forward(self, x):
        x = x[:, None]
        freqs = x * self.weights[None, :] * 2 * math.pi
        return torch.cat((x, freqs.sin(), freqs.cos()), dim=-1)


class NumberEmbedder(nn.Module):
    def __init__(self, features: int, dim: int = 256):
        super().__init__()
        self.features = features
        self.embedding = nn.Sequential(LearnedPositionalEmbedding(dim), nn.Linear(dim + 1, features))

    def forward(self, x):
        if not torch.is_tensor(x):
            x = torch.tensor(x, device=next(self.embedding.parameters()).device)
        shape = x.shape
        return self.embedding(x.reshape(-1)).view(*shape, self.features)


class NumberConditioner(Conditioner):
    """floats -> clamp to [min_val, max_val] -> normalise to [0, 1] -> NumberEmbedder -> [B, 1, dim]."""

    def __init__(self, output_dim: int, min_val: float = 0, max_val: float = 1):
        super().__init__(output_dim, output_dim)
        self.min_val, self.max_val = min_val, max_val
        self.embedder = NumberEmbedder(features=output_dim)
        self.device = next(self.embedder.parameters()).device

    def set_device(self, device):
        self.to(device)
        self.device = device

    def forward(self, floats: tp.List[float]):
        p = next(self.embedder.parameters())
        self.device = p.device
        x = torch.tensor([float(v) for v in floats]).to(self.device).clamp(self.min_val, self.max_val)
        x = ((x - self.min_val) / (self.max_val - self.min_val)).to(p.dtype)
        emb = self.embedder(x).unsqueeze(1)
        return [emb, torch.ones(emb.shape[0], 1).to(self.device)]


class IntConditioner(Conditioner):
    def __init__(self, output_dim: int, min_val: int = 0, max_val: int = 512):
        super().__init__(output_dim, output_dim)
        self.min_val, self.max_val = min_val, max_val
        self.int_embedder = nn.Embedding(max_val - min_val + 1, output_dim).requires_grad_(True)
        self.device = next(self.int_embedder.parameters()).device

    def set_device(self, device):
        self.to(device)
        self.device = device

    def forward(self, ints: tp.List[int]):
        self.device = next(self.int_embedder.parameters()).device
        idx = torch.tensor(ints).to(self.device).clamp(self.min_val, self.max_val)
        emb = self.int_embedder(idx).unsqueeze(1)
        return [emb, torch.ones(emb.shape[0], 1).to(self.device)]


class T5Conditioner(Conditioner):
    T5_MODEL_DIMS = {"t5-small": 512, "t5-base": 768, "t5-large": 1024, "t5-3b": 1024, "t5-11b": 1024,
                     "google/flan-t5-small": 512, "google/flan-t5-base": 768, "google/flan-t5-large": 1024,
                     "google/flan-t5-xl": 2048, "google/flan-t5-xxl": 4096}

    def __init__(self, output_dim: int, t5_model_name: str = "t5-base", max_length: int = 128,
                 enable_grad: bool = False, project_out: bool = False):
        assert t5_model_name in self.T5_MODEL_DIMS, f"Unknown T5 model name: {t5_model_name}"
        super().__init__(self.T5_MODEL_DIMS[t5_model_name], output_dim, project_out=project_out)
        from transformers import AutoTokenizer, T5EncoderModel
        self.max_length, self.enable_grad, self.device = max_length, enable_grad, "cpu"
        prev = logging.root.manager.disable
        logging.disable(logging.ERROR)
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self.tokenizer = AutoTokenizer.from_pretrained(t5_model_name)
                model = T5EncoderModel.from_pretrained(t5_model_name).train(enable_grad).requires_grad_(enable_grad)
                model = model.to(torch.float16)
        finally:
            logging.disable(prev)
        if enable_grad:
            self.model = model
        else:
            self.__dict__["model"] = model     # frozen: kept out of the state dict like the reference

    def set_device(self, device):
        self.to(device)
        self.model.to(device)
        self.device = device

    def forward(self, texts: tp.List[str]):
        enc = self.tokenizer(texts, truncation=True, max_length=self.max_length, padding="max_length",
                             return_tensors="pt")
        ids = enc["input_ids"].to(self.device)
        mask = enc["attention_mask"].to(self.device).to(torch.bool)
        self.model.eval()
        with torch.set_grad_enabled(self.enable_grad):
            emb = self.model(input_ids=ids, attention_mask=mask)["last_hidden_state"]
        emb = self.proj_out(emb.float()) * mask.unsqueeze(-1).float()
        return emb, mask


class MultiConditioner(nn.Module):
    """Applies one conditioner per key of the per-item metadata dicts."""

    def __init__(self, conditioners: tp.Dict[str, Conditioner], default_keys: tp.Dict[str, str] = {}):
        super().__init__()
        self.conditioners = nn.ModuleDict(conditioners)
        self.default_keys = default_keys

    def set_device(self, device):
        for m in self.conditioners.values():
            m.set_device(device)

    def forward(self, batch_metadata: tp.List[tp.Dict[str, tp.Any]]):
        out = {}
        for key, cond in self.conditioners.items():
            ck, inputs = key, []
            for item in batch_metadata:
                if ck not in item:
                    if ck in self.default_keys:
                        ck = self.default_keys[ck]
                    else:
                        raise ValueError(f"Conditioner key {ck} not found in batch metadata")
                v = item[ck]
                inputs.append(v[0] if isinstance(v, (list, tuple)) and len(v) == 1 else v)
            out[key] = cond(inputs)
        return out


_TYPES = {"t5": T5Conditioner, "number": NumberConditioner, "int": IntConditioner}


def create_multi_conditioner_from_conditioning_config(config: tp.Dict[str, tp.Any]) -> MultiConditioner:
    conditioners = {}
    for info in config["configs"]:
        kind = info["type"]
        if kind not in _TYPES:
            raise NotImplementedError(f"conditioner type '{kind}' is outside this build's scope "
                                      f"(available: {sorted(_TYPES)})")
        conditioners[info["id"]] = _TYPES[kind](**{"output_dim": config["cond_dim"], **info["config"]})
    return MultiConditioner(conditioners, default_keys=config.get("default_keys", {}))
