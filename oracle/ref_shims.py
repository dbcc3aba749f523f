"""sys.modules shims that let the real reference import in the build container.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference imports several
third-party packages at module import time that are absent offline
(``dac``, ``alias_free_torch``, ``x_transformers``, ``einops_exts``,
``vector_quantize_pytorch``, ``k_diffusion``).  On the hot path only three of
their symbols are ever *called*:

* ``dac.nn.layers.WNConv1d``            = ``weight_norm(nn.Conv1d(...))``
* ``dac.nn.layers.WNConvTranspose1d``   = ``weight_norm(nn.ConvTranspose1d(...))``
  (descript-audio-codec 1.0.0, ``dac/nn/layers.py``; pinned by reference
  ``setup.py:12``)
* ``k_diffusion`` sampler entry points  (restated in ``sampler_oracle``)

Everything else is a never-instantiated stub.  ``/root/reference`` exists only
in the build container; on the GPU box ``reference_available()`` is False and
tests that need it are skipped (the committed golden vectors stand in).
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "stable_audio_tools"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shims():
    """Install the third-party stubs; idempotent."""
    import torch
    from torch import nn
    from torch.nn.utils import weight_norm

    if "dac.nn.layers" in sys.modules and getattr(sys.modules["dac.nn.layers"], "_satb_shim", False):
        return

    class _Stub(nn.Module):
        def __init__(self, *a, **k):
            raise RuntimeError("stub: not on the hot path")

    class Snake1d(nn.Module):  # dac 1.0.0 semantics; only reached by out-of-scope blocks
        def __init__(self, c):
            super().__init__()
            self.alpha = nn.Parameter(torch.ones(1, c, 1))

        def forward(self, x):
            return x + (self.alpha + 1e-9).reciprocal() * torch.sin(self.alpha * x).pow(2)

    _mod("dac")
    _mod("dac.nn")
    _mod("dac.nn.layers",
         WNConv1d=lambda *a, **k: weight_norm(nn.Conv1d(*a, **k)),
         WNConvTranspose1d=lambda *a, **k: weight_norm(nn.ConvTranspose1d(*a, **k)),
         Snake1d=Snake1d, _satb_shim=True)
    _mod("dac.nn.quantize", ResidualVectorQuantize=_Stub)
    _mod("alias_free_torch", Activation1d=_Stub)
    _mod("x_transformers", ContinuousTransformerWrapper=_Stub, Encoder=_Stub, Decoder=_Stub)

    def rearrange_many(ts, p, **k):
        import einops
        return [einops.rearrange(t, p, **k) for t in ts]

    _mod("einops_exts", rearrange_many=rearrange_many)
    _mod("vector_quantize_pytorch", ResidualVQ=_Stub, FSQ=_Stub)
    K = _mod("k_diffusion")
    K.external = _mod("k_diffusion.external")
    K.sampling = _mod("k_diffusion.sampling")
    K.utils = _mod("k_diffusion.utils")
    # the restated sampler pieces (third-party arithmetic, parity unpinned)
    from . import sampler_oracle as so
    K.external.VDenoiser = so.VDenoiser
    K.sampling.get_sigmas_polyexponential = so.get_sigmas_polyexponential
    K.sampling.sample_dpmpp_2m_sde = so.sample_dpmpp_2m_sde
    K.sampling.sample_dpmpp_3m_sde = so.sample_dpmpp_3m_sde
    K.utils.append_dims = so.append_dims


def import_reference():
    """Import the real reference package (build container only).

    Returns the ``stable_audio_tools`` module object of the reference.  The
    repo's own drop-in package has the same top-level name, so the reference is
    imported under a private ``sys.modules`` snapshot and handed back as a
    namespace; callers use attributes of the returned object only.
    """
    if not reference_available():
        raise RuntimeError("/root/reference is not available on this machine")
    install_shims()
    saved = {k: v for k, v in sys.modules.items() if k == "stable_audio_tools" or k.startswith("stable_audio_tools.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import importlib
        ref = importlib.import_module("stable_audio_tools")
        importlib.import_module("stable_audio_tools.models.dit")
        importlib.import_module("stable_audio_tools.models.transformer")
        importlib.import_module("stable_audio_tools.models.autoencoders")
        importlib.import_module("stable_audio_tools.models.diffusion")
        importlib.import_module("stable_audio_tools.models.pretransforms")
        importlib.import_module("stable_audio_tools.models.factory")
        importlib.import_module("stable_audio_tools.models.bottleneck")
        importlib.import_module("stable_audio_tools.models.blocks")
        importlib.import_module("stable_audio_tools.inference.generation")
        importlib.import_module("stable_audio_tools.inference.sampling")
        ref_modules = {k: v for k, v in sys.modules.items()
                       if k == "stable_audio_tools" or k.startswith("stable_audio_tools.")}
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in list(sys.modules):
            if k == "stable_audio_tools" or k.startswith("stable_audio_tools."):
                del sys.modules[k]
        sys.modules.update(saved)
    ns = types.SimpleNamespace()
    ns.root = ref
    ns.modules = ref_modules
    ns.dit = ref_modules["stable_audio_tools.models.dit"]
    ns.transformer = ref_modules["stable_audio_tools.models.transformer"]
    ns.autoencoders = ref_modules["stable_audio_tools.models.autoencoders"]
    ns.diffusion = ref_modules["stable_audio_tools.models.diffusion"]
    ns.pretransforms = ref_modules["stable_audio_tools.models.pretransforms"]
    ns.factory = ref_modules["stable_audio_tools.models.factory"]
    ns.bottleneck = ref_modules["stable_audio_tools.models.bottleneck"]
    ns.blocks = ref_modules["stable_audio_tools.models.blocks"]
    ns.generation = ref_modules["stable_audio_tools.inference.generation"]
    ns.sampling = ref_modules["stable_audio_tools.inference.sampling"]
    return ns


class reference_modules:
    """with reference_modules(ref): ... -> the reference package is what `import stable_audio_tools...` resolves to
    (its factories import their siblings lazily, e.g. models/factory.py:7-9); the drop-in package's modules are
    restored on exit."""

    def __init__(self, ref):
        self.ref = ref

    @staticmethod
    def _ours():
        return {k: v for k, v in sys.modules.items() if k == "stable_audio_tools" or k.startswith("stable_audio_tools.")}

    def __enter__(self):
        self.saved = self._ours()
        for k in self.saved:
            del sys.modules[k]
        sys.modules.update(self.ref.modules)
        sys.path.insert(0, REFERENCE_ROOT)
        return self

    def __exit__(self, *exc):
        sys.path.remove(REFERENCE_ROOT)
        self.ref.modules.update(self._ours())      # keep lazily imported reference modules for the next use
        for k in list(sys.modules):
            if k == "stable_audio_tools" or k.startswith("stable_audio_tools."):
                del sys.modules[k]
        sys.modules.update(self.saved)
