"""Samplers for the v-objective denoiser.

``sample_k`` keeps the reference signature and behaviour (``inference/sampling.py:144-228``):
polyexponential sigma schedule, initial noise scaled by sigma_0, variation / inpainting
initialisation and the inpainting callback, sampler dispatch by name.  The k-diffusion 0.1.1
pieces it relies on (``VDenoiser``, ``get_sigmas_polyexponential``, DPM-Solver++(2M/3M) SDE)
are an un-vendored third-party dependency of the reference (``setup.py:21``) that is absent
offline, so they are restated here from the published algorithms (Karras et al. 2022;
Lu et al. 2022) - parity for those is unpinned by the reference.
k-diffusion's Brownian-tree noise (torchsde) is replaced by one ``randn_like`` draw per
step, which has the same distribution over the disjoint sigma intervals; pass
``noise_sampler=`` to inject an explicit sequence.
"""
import math

import torch


def exists(x):
    return x is not None


def append_dims(x, target_dims):
    extra = target_dims - x.ndim
    if extra < 0:
        raise ValueError("input has more dims than the target")
    return x[(...,) + (None,) * extra]


def get_sigmas_polyexponential(n, sigma_min, sigma_max, rho=1.0, device="cpu"):
    ramp = torch.linspace(1, 0, n, device=device) ** rho
    lo, hi = math.log(sigma_min), math.log(sigma_max)
    sigmas = torch.exp(ramp * (hi - lo) + lo)
    return torch.cat([sigmas, sigmas.new_zeros([1])])


class VDenoiser(torch.nn.Module):
    """D(x, sigma) = F(x c_in, atan(sigma) 2/pi) c_out + x c_skip with sigma_data = 1."""

    def __init__(self, inner_model):
        super().__init__()
        self.inner_model = inner_model
        self.sigma_data = 1.0

    def get_scalings(self, sigma):
        denom = sigma ** 2 + self.sigma_data ** 2
        return self.sigma_data ** 2 / denom, -sigma * self.sigma_data / denom ** 0.5, 1 / denom ** 0.5

    def sigma_to_t(self, sigma):
        return sigma.atan() / math.pi * 2

    def forward(self, input, sigma, **kwargs):
        c_skip, c_out, c_in = (append_dims(c, input.ndim) for c in self.get_scalings(sigma))
        return self.inner_model(input * c_in, self.sigma_to_t(sigma), **kwargs) * c_out + input * c_skip


def _noise_fn(x, noise_sampler):
    if noise_sampler is not None:
        return noise_sampler
    return lambda sigma, sigma_next: torch.randn_like(x)


@torch.no_grad()
def sample_dpmpp_2m_sde(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1.0, s_noise=1.0,
                        noise_sampler=None, solver_type="midpoint"):
    noise = _noise_fn(x, noise_sampler)
    extra_args = extra_args or {}
    ones = x.new_ones([x.shape[0]])
    sig = [float(s) for s in sigmas]          # host copies: no device sync inside the loop
    prev_den, prev_h = None, None
    for i in range(len(sig) - 1):
        den = model(x, sigmas[i] * ones, **extra_args)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": den})
        h = None
        if sig[i + 1] == 0:
            x = den
        else:
            h = math.log(sig[i]) - math.log(sig[i + 1])
            eta_h = eta * h
            x = sig[i + 1] / sig[i] * math.exp(-eta_h) * x + (-math.expm1(-h - eta_h)) * den
            if prev_den is not None:
                r = prev_h / h
                if solver_type == "heun":
                    x = x + ((-math.expm1(-h - eta_h)) / (-h - eta_h) + 1) * (1 / r) * (den - prev_den)
                else:
                    x = x + 0.5 * (-math.expm1(-h - eta_h)) * (1 / r) * (den - prev_den)
            if eta:
                x = x + noise(sigmas[i], sigmas[i + 1]) * (sig[i + 1] * math.sqrt(-math.expm1(-2 * eta_h)) * s_noise)
        prev_den, prev_h = den, h
    return x


@torch.no_grad()
def sample_dpmpp_3m_sde(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1.0, s_noise=1.0,
                        noise_sampler=None):
    noise = _noise_fn(x, noise_sampler)
    extra_args = extra_args or {}
    ones = x.new_ones([x.shape[0]])
    sig = [float(s) for s in sigmas]
    den_1 = den_2 = None
    h_1 = h_2 = None
    for i in range(len(sig) - 1):
        den = model(x, sigmas[i] * ones, **extra_args)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": den})
        h = None
        if sig[i + 1] == 0:
            x = den
        else:
            h = math.log(sig[i]) - math.log(sig[i + 1])
            h_eta = h * (eta + 1)
            x = math.exp(-h_eta) * x + (-math.expm1(-h_eta)) * den
            if h_2 is not None:
                r0, r1 = h_1 / h, h_2 / h
                d1_0 = (den - den_1) / r0
                d1_1 = (den_1 - den_2) / r1
                d1 = d1_0 + (d1_0 - d1_1) * (r0 / (r0 + r1))
                d2 = (d1_0 - d1_1) / (r0 + r1)
                phi_2 = math.expm1(-h_eta) / h_eta + 1
                phi_3 = phi_2 / h_eta - 0.5
                x = x + phi_2 * d1 - phi_3 * d2
            elif h_1 is not None:
                r = h_1 / h
                phi_2 = math.expm1(-h_eta) / h_eta + 1
                x = x + phi_2 * ((den - den_1) / r)
            if eta:
                x = x + noise(sigmas[i], sigmas[i + 1]) * (sig[i + 1] * math.sqrt(-math.expm1(-2 * h * eta)) * s_noise)
        den_1, den_2 = den, den_1
        h_1, h_2 = h, h_1
    return x


SAMPLERS = {"dpmpp-2m-sde": sample_dpmpp_2m_sde, "dpmpp-3m-sde": sample_dpmpp_3m_sde}


def get_bmask(i, steps, mask):
    """Shrinking hard mask for soft-mask inpainting (reference generation.py:277-281)."""
    return torch.where(mask <= (i + 1) / steps, 1, 0)


def sample_k(model_fn, noise, init_data=None, mask=None, steps=100, sampler_type="dpmpp-2m-sde", sigma_min=0.5,
             sigma_max=50, rho=1.0, device="cuda", callback=None, cond_fn=None, disable_tqdm: bool = False,
             noise_sampler=None, **extra_args):
    if cond_fn is not None:
        raise NotImplementedError("guidance through cond_fn needs autograd through the denoiser (inference-only here)")
    if sampler_type not in SAMPLERS:
        raise NotImplementedError(f"sampler '{sampler_type}' is not restated; available: {sorted(SAMPLERS)}")
    denoiser = VDenoiser(model_fn)
    sigmas = get_sigmas_polyexponential(steps, sigma_min, sigma_max, rho, device=device)
    noise = noise * sigmas[0]
    wrapped_callback = callback
    if mask is None and exists(init_data):
        x = init_data + noise                                   # variation
    elif exists(mask) and exists(init_data):
        bmask = get_bmask(0, steps, mask)                       # inpainting
        x = (init_data + noise) * bmask + noise * (1 - bmask)

        def inpainting_callback(args):
            i, xx, sigma = args["i"], args["x"], args["sigma"]
            noised = init_data + torch.randn_like(init_data) * sigma
            bm = get_bmask(i, steps, mask)
            xx[:, :, :] = (noised * bm + xx * (1 - bm))[:, :, :]

        if callback is None:
            wrapped_callback = inpainting_callback
        else:
            def wrapped_callback(args):
                return inpainting_callback(args), callback(args)
    else:
        x = noise
    return SAMPLERS[sampler_type](denoiser, x, sigmas, disable=disable_tqdm, callback=wrapped_callback,
                                  extra_args=extra_args, noise_sampler=noise_sampler)
