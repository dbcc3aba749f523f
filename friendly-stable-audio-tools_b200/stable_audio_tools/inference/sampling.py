"""Samplers for the v-objective denoiser.

``sample_k`` keeps the reference signature and behaviour (``inference/sampling.py:144-228``):
polyexponential sigma schedule, initial noise scaled by sigma_0, variation / inpainting
initialisation and the inpainting callback, sampler dispatch by name.  The k-diffusion 0.1.1
pieces it relies on (``VDenoiser``, ``get_sigmas_polyexponential``, DPM-Solver++(2M/3M) SDE, and
the ``k-*`` samplers: Heun, linear multistep, DPM-2, DPM-Solver++(2S) ancestral, DPM-Solver
fast / adaptive) are an un-vendored third-party dependency of the reference (``setup.py:21``) that
is absent offline, so they are restated here from the published algorithms (Karras et al. 2022,
Alg. 1/2; Lu et al. 2022, DPM-Solver and DPM-Solver++) - parity for those is unpinned by the
reference; ``tests/test_host_logic.py`` checks every one of them against the
closed-form probability-flow solution of a Gaussian toy problem.
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


class MultistepSdeStepper:
    """DPM-Solver++(2M) SDE / DPM-Solver++(3M) SDE as a state machine with one model call per ``step()``.

    Every update of both samplers is linear in the tensors involved,
        x_next = A x + B den + C den_1 + D den_2 + S noise,
    with scalars that depend only on the sigma schedule (``coeffs``).  On CUDA, with the standard ``VDenoiser``
    wrapper and no callback, the denoiser scalings, this update and the scaling of the next model input run
    as ONE kernel (``satb_sampler_update``) instead of ~20 elementwise launches; otherwise the same algebra is
    evaluated with torch ops (CPU tensors in the tests, callbacks that need ``denoised`` before the update).
    """

    def __init__(self, model, x, sigmas, order=3, extra_args=None, callback=None, eta=1.0, s_noise=1.0,
                 noise_sampler=None, solver_type="midpoint"):
        self.model, self.x, self.sigmas, self.order = model, x, sigmas, order
        self.extra_args, self.callback = extra_args or {}, callback
        self.eta, self.s_noise, self.solver_type = eta, s_noise, solver_type
        self.noise = _noise_fn(x, noise_sampler)
        self.sig = [float(v) for v in sigmas]          # host copies: no device sync inside the loop
        self.ones = x.new_ones([x.shape[0]])
        self.i = 0
        self.den_1 = self.den_2 = None
        self.h_1 = self.h_2 = None
        self.fused = (isinstance(model, VDenoiser) and callback is None and x.is_cuda and x.dtype == torch.float32
                      and x.numel() % 4 == 0)
        # the native DiT (directly, or as DiTWrapper.model) can run a call as one CUDA-graph launch; its output is then a
        # static buffer, which is safe here because the fused update consumes v before the next model call
        self.graph_dit = None
        if self.fused:
            inner = getattr(model, "inner_model", None)
            for cand in (inner, getattr(inner, "model", None)):
                if cand is not None and hasattr(cand, "cuda_graph") and hasattr(cand, "_graph_forward"):
                    self.graph_dit = cand
        self.x_in = None                               # x * c_in(sigma_i), produced by the previous fused update

    def coeffs(self, i):
        """(A, B, C, D, S, h) of step i; C / D are 0 while the history is shorter than the order."""
        sig, eta = self.sig, self.eta
        if sig[i + 1] == 0:
            return 0.0, 1.0, 0.0, 0.0, 0.0, None
        h = math.log(sig[i]) - math.log(sig[i + 1])
        C = D = 0.0
        if self.order == 2:
            eta_h = eta * h
            A = sig[i + 1] / sig[i] * math.exp(-eta_h)
            E = -math.expm1(-h - eta_h)
            B = E
            if self.den_1 is not None:
                r = self.h_1 / h
                K = (E / (-h - eta_h) + 1) / r if self.solver_type == "heun" else 0.5 * E / r
                B, C = B + K, -K
            S = sig[i + 1] * math.sqrt(-math.expm1(-2 * eta_h)) * self.s_noise if eta else 0.0
        else:
            h_eta = h * (eta + 1)
            A = math.exp(-h_eta)
            B = -math.expm1(-h_eta)
            phi_2 = math.expm1(-h_eta) / h_eta + 1
            if self.h_2 is not None:
                r0, r1 = self.h_1 / h, self.h_2 / h
                w, q = r0 / (r0 + r1), 1.0 / (r0 + r1)
                phi_3 = phi_2 / h_eta - 0.5
                P, Q = phi_2 * (1 + w) - phi_3 * q, -phi_2 * w + phi_3 * q
                B, C, D = B + P / r0, -P / r0 + Q / r1, -Q / r1
            elif self.h_1 is not None:
                r = self.h_1 / h
                B, C = B + phi_2 / r, -phi_2 / r
            S = sig[i + 1] * math.sqrt(-math.expm1(-2 * h * eta)) * self.s_noise if eta else 0.0
        return A, B, C, D, S, h

    def step(self, i=None):
        """One model evaluation + update at schedule index i (default: the next one); returns the new x."""
        i = self.i if i is None else i
        x, sig = self.x, self.sig
        A, B, C, D, S, h = self.coeffs(i)
        nz = self.noise(self.sigmas[i], self.sigmas[i + 1]) if S != 0.0 else None
        if self.fused:
            from .. import _native
            c_skip, c_out, c_in = (float(c) for c in self.model.get_scalings(torch.tensor(sig[i], dtype=torch.float64)))
            if self.x_in is None:
                self.x_in = x * c_in
            t = self.model.sigma_to_t(self.sigmas[i]) * self.ones
            v = self.model.inner_model(self.x_in, t, **self.extra_args)
            v = v.to(torch.float32).contiguous()
            c_in_next = 1.0 / math.sqrt(sig[i + 1] ** 2 + self.model.sigma_data ** 2)
            den, x_next, x_in_next = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
            xc = x.contiguous()
            _native.check(_native.lib().satb_sampler_update(
                _native.ptr(xc), _native.ptr(v), _native.ptr(self.den_1 if C != 0.0 else None),
                _native.ptr(self.den_2 if D != 0.0 else None), _native.ptr(nz.contiguous() if nz is not None else None),
                _native.ptr(den), _native.ptr(x_next), _native.ptr(x_in_next), x.numel(), c_skip, c_out, A, B, C, D, S,
                c_in_next, _native.stream_ptr(x.device)))
            self.x_in = x_in_next
        else:
            den = self.model(x, self.sigmas[i] * self.ones, **self.extra_args)
            if self.callback is not None:
                self.callback({"x": x, "i": i, "sigma": self.sigmas[i], "sigma_hat": self.sigmas[i], "denoised": den})
            x_next = den if (A == 0.0 and B == 1.0) else A * x + B * den
            if C != 0.0:
                x_next = x_next + C * self.den_1
            if D != 0.0:
                x_next = x_next + D * self.den_2
            if nz is not None:
                x_next = x_next + S * nz
        self.den_1, self.den_2 = den, self.den_1
        self.h_1, self.h_2 = h, self.h_1
        self.x = x_next
        self.i = i + 1
        return x_next

    def run(self):
        prev = None
        if self.graph_dit is not None:
            prev, self.graph_dit.cuda_graph = self.graph_dit.cuda_graph, True
        try:
            for _ in range(len(self.sig) - 1):
                self.step()
        finally:
            if self.graph_dit is not None:
                self.graph_dit.cuda_graph = prev
        return self.x


@torch.no_grad()
def sample_dpmpp_2m_sde(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1.0, s_noise=1.0,
                        noise_sampler=None, solver_type="midpoint"):
    """DPM-Solver++(2M) SDE (Lu et al. 2022; k-diffusion's sample_dpmpp_2m_sde, eta = 1, 'midpoint')."""
    return MultistepSdeStepper(model, x, sigmas, 2, extra_args, callback, eta, s_noise, noise_sampler, solver_type).run()


@torch.no_grad()
def sample_dpmpp_3m_sde(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1.0, s_noise=1.0,
                        noise_sampler=None):
    """DPM-Solver++(3M) SDE (k-diffusion's sample_dpmpp_3m_sde, eta = 1)."""
    return MultistepSdeStepper(model, x, sigmas, 3, extra_args, callback, eta, s_noise, noise_sampler).run()


def _to_d(x, sigma, denoised):
    """Karras ODE derivative dx/dsigma = (x - D(x, sigma)) / sigma."""
    return (x - denoised) / sigma


@torch.no_grad()
def sample_heun(model, x, sigmas, extra_args=None, callback=None, disable=None, **_):
    """Karras et al. 2022, Algorithm 1 without churn: Euler step + trapezoidal correction."""
    extra_args = extra_args or {}
    ones = x.new_ones([x.shape[0]])
    sig = [float(v) for v in sigmas]
    for i in range(len(sig) - 1):
        den = model(x, sigmas[i] * ones, **extra_args)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": den})
        d = _to_d(x, sig[i], den)
        dt = sig[i + 1] - sig[i]
        if sig[i + 1] == 0:
            x = x + d * dt
        else:
            x_2 = x + d * dt
            den_2 = model(x_2, sigmas[i + 1] * ones, **extra_args)
            x = x + (d + _to_d(x_2, sig[i + 1], den_2)) * (0.5 * dt)
    return x


@torch.no_grad()
def sample_dpm_2(model, x, sigmas, extra_args=None, callback=None, disable=None, **_):
    """Second-order sampler with the midpoint taken in log-sigma (Karras et al. 2022, Algorithm 2 without churn)."""
    extra_args = extra_args or {}
    ones = x.new_ones([x.shape[0]])
    sig = [float(v) for v in sigmas]
    for i in range(len(sig) - 1):
        den = model(x, sigmas[i] * ones, **extra_args)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": den})
        d = _to_d(x, sig[i], den)
        if sig[i + 1] == 0:
            x = x + d * (sig[i + 1] - sig[i])
        else:
            sigma_mid = math.exp(0.5 * (math.log(sig[i]) + math.log(sig[i + 1])))
            x_2 = x + d * (sigma_mid - sig[i])
            den_2 = model(x_2, sigma_mid * ones, **extra_args)
            x = x + _to_d(x_2, sigma_mid, den_2) * (sig[i + 1] - sig[i])
    return x


def _lms_coeff(order, t, i, j):
    """Integral over [t_i, t_{i+1}] of the j-th Lagrange basis polynomial through t_i, t_{i-1}, ..., t_{i-order+1}
    (exact polynomial integration)."""
    import numpy as np
    poly = np.poly1d([1.0])
    for k in range(order):
        if k != j:
            poly = poly * np.poly1d([1.0, -t[i - k]]) / (t[i - j] - t[i - k])
    integ = poly.integ()
    return float(integ(t[i + 1]) - integ(t[i]))


@torch.no_grad()
def sample_lms(model, x, sigmas, extra_args=None, callback=None, disable=None, order=4, **_):
    """Linear multistep (Adams-Bashforth in sigma) of order <= 4 over the derivative history."""
    extra_args = extra_args or {}
    ones = x.new_ones([x.shape[0]])
    sig = [float(v) for v in sigmas]
    ds = []
    for i in range(len(sig) - 1):
        den = model(x, sigmas[i] * ones, **extra_args)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": den})
        ds.append(_to_d(x, sig[i], den))
        if len(ds) > order:
            ds.pop(0)
        cur = min(i + 1, order)
        for j, d in zip(range(cur), reversed(ds)):
            x = x + d * _lms_coeff(cur, sig, i, j)
    return x


def get_ancestral_step(sigma_from, sigma_to, eta=1.0):
    """Split a step into a deterministic part down to sigma_down and fresh noise of scale sigma_up."""
    if not eta:
        return sigma_to, 0.0
    sigma_up = min(sigma_to, eta * math.sqrt(sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2))
    return math.sqrt(sigma_to ** 2 - sigma_up ** 2), sigma_up


@torch.no_grad()
def sample_dpmpp_2s_ancestral(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1.0, s_noise=1.0,
                              noise_sampler=None):
    """DPM-Solver++(2S) with ancestral noise (Lu et al. 2022, data-prediction, single-step second order)."""
    noise = _noise_fn(x, noise_sampler)
    extra_args = extra_args or {}
    ones = x.new_ones([x.shape[0]])
    sig = [float(v) for v in sigmas]
    for i in range(len(sig) - 1):
        den = model(x, sigmas[i] * ones, **extra_args)
        sigma_down, sigma_up = get_ancestral_step(sig[i], sig[i + 1], eta)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": den})
        if sigma_down == 0:
            x = x + _to_d(x, sig[i], den) * (sigma_down - sig[i])
        else:
            t, t_next = -math.log(sig[i]), -math.log(sigma_down)
            h = t_next - t
            s_mid = t + 0.5 * h
            x_2 = (math.exp(-s_mid) / math.exp(-t)) * x - math.expm1(-0.5 * h) * den
            den_2 = model(x_2, math.exp(-s_mid) * ones, **extra_args)
            x = (math.exp(-t_next) / math.exp(-t)) * x - math.expm1(-h) * den_2
        if sig[i + 1] > 0 and sigma_up > 0:
            x = x + noise(sigmas[i], sigmas[i + 1]) * (s_noise * sigma_up)
    return x


class _DPMSolver:
    """DPM-Solver (Lu et al. 2022) in t = -log(sigma) with noise prediction eps = (x - D(x, sigma)) / sigma:
    singlestep orders 1-3, the fixed-budget schedule ("fast") and the adaptive order-2/3 pair with a
    PID step-size controller ("adaptive")."""

    def __init__(self, model, extra_args=None, callback=None):
        self.model, self.extra_args, self.callback = model, extra_args or {}, callback
        self.nfe = 0

    @staticmethod
    def sigma(t):
        return math.exp(-t)

    def eps(self, cache, key, x, t):
        if key in cache:
            return cache[key], cache
        sig = self.sigma(t)
        eps = (x - self.model(x, x.new_ones([x.shape[0]]) * sig, **self.extra_args)) / sig
        self.nfe += 1
        return eps, {key: eps, **cache}

    def step1(self, x, t, t_next, cache=None):
        cache = cache or {}
        h = t_next - t
        eps, cache = self.eps(cache, "eps", x, t)
        return x - self.sigma(t_next) * math.expm1(h) * eps, cache

    def step2(self, x, t, t_next, r1=0.5, cache=None):
        cache = cache or {}
        h = t_next - t
        eps, cache = self.eps(cache, "eps", x, t)
        s1 = t + r1 * h
        u1 = x - self.sigma(s1) * math.expm1(r1 * h) * eps
        eps_r1, cache = self.eps(cache, "eps_r1", u1, s1)
        x_2 = x - self.sigma(t_next) * math.expm1(h) * eps - self.sigma(t_next) / (2 * r1) * math.expm1(h) * (eps_r1 - eps)
        return x_2, cache

    def step3(self, x, t, t_next, r1=1 / 3, r2=2 / 3, cache=None):
        cache = cache or {}
        h = t_next - t
        eps, cache = self.eps(cache, "eps", x, t)
        s1, s2 = t + r1 * h, t + r2 * h
        u1 = x - self.sigma(s1) * math.expm1(r1 * h) * eps
        eps_r1, cache = self.eps(cache, "eps_r1", u1, s1)
        u2 = (x - self.sigma(s2) * math.expm1(r2 * h) * eps
              - self.sigma(s2) * (r2 / r1) * (math.expm1(r2 * h) / (r2 * h) - 1) * (eps_r1 - eps))
        eps_r2, cache = self.eps(cache, "eps_r2", u2, s2)
        x_3 = (x - self.sigma(t_next) * math.expm1(h) * eps
               - self.sigma(t_next) / r2 * (math.expm1(h) / h - 1) * (eps_r2 - eps))
        return x_3, cache

    def _report(self, x, i, t, cache):
        if self.callback is not None and "eps" in cache:
            sig = self.sigma(t)
            self.callback({"x": x, "i": i, "t": t, "sigma": x.new_tensor(sig), "sigma_hat": x.new_tensor(sig),
                           "denoised": x - sig * cache["eps"]})

    def fast(self, x, t_start, t_end, nfe):
        if nfe < 1:
            raise ValueError("nfe must be at least 1")
        m = nfe // 3 + 1
        ts = [t_start + (t_end - t_start) * k / m for k in range(m + 1)]
        orders = [3] * (m - 2) + [2, 1] if nfe % 3 == 0 else [3] * (m - 1) + [nfe % 3]
        for i, order in enumerate(orders):
            t, t_next = ts[i], ts[i + 1]
            eps, cache = self.eps({}, "eps", x, t)
            self._report(x, i, t, cache)
            if order == 1:
                x, _ = self.step1(x, t, t_next, cache=cache)
            elif order == 2:
                x, _ = self.step2(x, t, t_next, cache=cache)
            else:
                x, _ = self.step3(x, t, t_next, cache=cache)
        return x

    def adaptive(self, x, t_start, t_end, order=3, rtol=0.05, atol=0.0078, h_init=0.05, pcoeff=0.0, icoeff=1.0,
                 dcoeff=0.0, accept_safety=0.81):
        if order not in (2, 3):
            raise ValueError("order should be 2 or 3")
        forward = t_end > t_start
        h = abs(h_init) * (1 if forward else -1)
        b1, b2, b3 = (pcoeff + icoeff + dcoeff) / order, -(pcoeff + 2 * dcoeff) / order, dcoeff / order
        errs = None
        s, x_prev, n_steps = t_start, x, 0
        while (s < t_end - 1e-5) if forward else (s > t_end + 1e-5):
            t = min(t_end, s + h) if forward else max(t_end, s + h)
            eps, cache = self.eps({}, "eps", x, s)
            denoised = x - self.sigma(s) * eps          # reported below: the estimate at the OLD (x, s)
            if order == 2:
                x_low, cache = self.step1(x, s, t, cache=cache)
                x_high, cache = self.step2(x, s, t, cache=cache)
            else:
                x_low, cache = self.step2(x, s, t, r1=1 / 3, cache=cache)
                x_high, cache = self.step3(x, s, t, cache=cache)
            delta = torch.maximum(torch.full_like(x_low, atol), rtol * torch.maximum(x_low.abs(), x_prev.abs()))
            error = float(torch.linalg.norm((x_low - x_high) / delta) / x.numel() ** 0.5)
            inv = 1.0 / (error + 1e-8)
            if errs is None:
                errs = [inv, inv, inv]
            errs[0] = inv
            factor = errs[0] ** b1 * errs[1] ** b2 * errs[2] ** b3
            factor = 1 + math.atan(factor - 1)
            accept = factor >= accept_safety
            if accept:
                errs[2], errs[1] = errs[1], errs[0]
                x_prev, x, s = x_low, x_high, t
            h *= factor
            n_steps += 1
            # k-diffusion's contract (the inpainting callback relies on it): the callback fires on EVERY iteration,
            # accepted or rejected, with i = the iteration count, x = the state after this iteration and sigma at
            # the proposed t
            if self.callback is not None:
                sig = x.new_tensor(self.sigma(t))
                self.callback({"x": x, "i": n_steps - 1, "t": t, "sigma": sig, "sigma_hat": sig, "denoised": denoised})
        return x


@torch.no_grad()
def sample_dpm_fast(model, x, sigma_min, sigma_max, n, extra_args=None, callback=None, disable=None, **_):
    """DPM-Solver with a fixed budget of n model evaluations between sigma_max and sigma_min."""
    if sigma_min <= 0 or sigma_max <= 0:
        raise ValueError("sigma_min and sigma_max must not be 0")
    return _DPMSolver(model, extra_args, callback).fast(x, -math.log(sigma_max), -math.log(sigma_min), n)


@torch.no_grad()
def sample_dpm_adaptive(model, x, sigma_min, sigma_max, extra_args=None, callback=None, disable=None, order=3,
                        rtol=0.05, atol=0.0078, h_init=0.05, pcoeff=0.0, icoeff=1.0, dcoeff=0.0, accept_safety=0.81, **_):
    """DPM-Solver-12/23 with adaptive step size (PID controller on the embedded error estimate)."""
    if sigma_min <= 0 or sigma_max <= 0:
        raise ValueError("sigma_min and sigma_max must not be 0")
    return _DPMSolver(model, extra_args, callback).adaptive(x, -math.log(sigma_max), -math.log(sigma_min), order, rtol,
                                                            atol, h_init, pcoeff, icoeff, dcoeff, accept_safety)


# sampler_type -> (function, takes a sigma schedule?)  (reference inference/sampling.py:211-228)
SAMPLERS = {
    "k-heun": sample_heun, "k-lms": sample_lms, "k-dpmpp-2s-ancestral": sample_dpmpp_2s_ancestral, "k-dpm-2": sample_dpm_2,
    "k-dpm-fast": sample_dpm_fast, "k-dpm-adaptive": sample_dpm_adaptive,
    "dpmpp-2m-sde": sample_dpmpp_2m_sde, "dpmpp-3m-sde": sample_dpmpp_3m_sde,
}


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
    if sampler_type == "k-dpm-fast":
        return sample_dpm_fast(denoiser, x, sigma_min, sigma_max, steps, disable=disable_tqdm, callback=wrapped_callback,
                               extra_args=extra_args)
    if sampler_type == "k-dpm-adaptive":
        return sample_dpm_adaptive(denoiser, x, sigma_min, sigma_max, rtol=0.01, atol=0.01, disable=disable_tqdm,
                                   callback=wrapped_callback, extra_args=extra_args)
    return SAMPLERS[sampler_type](denoiser, x, sigmas, disable=disable_tqdm, callback=wrapped_callback,
                                  extra_args=extra_args, noise_sampler=noise_sampler)


@torch.no_grad()
def sample_discrete_euler(model, x, steps, sigma_max=1, callback=None, **extra_args):
    """Rectified-flow sampling (reference inference/sampling.py:29-60): the network predicts the velocity and
    the state is integrated from t = sigma_max down to 0 on a uniform grid, x += (t_next - t) * v(x, t)."""
    ts = torch.linspace(sigma_max, 0, steps + 1)
    ones = x.new_ones([x.shape[0]])
    for i in range(steps):
        t_curr, t_next = float(ts[i]), float(ts[i + 1])
        v = model(x, t_curr * ones, **extra_args)
        if callback is not None:
            callback({"x": x, "i": i, "t": t_curr, "denoised": x - t_curr * v})
        x = x + (t_next - t_curr) * v
    return x


def sample_rf(model_fn, noise, init_data=None, steps=100, sigma_max=1, device="cuda", callback=None, cond_fn=None,
              disable_tqdm: bool = False, **extra_args):
    """Reference inference/sampling.py:236-270: plain noise, or a variation that starts from the
    (1 - sigma_max, sigma_max) interpolation of init_data and noise."""
    if cond_fn is not None:
        raise NotImplementedError("guidance through cond_fn needs autograd through the model (inference-only here)")
    sigma_max = min(sigma_max, 1)
    x = noise if init_data is None else init_data * (1 - sigma_max) + noise * sigma_max
    return sample_discrete_euler(model_fn, x, steps, sigma_max, callback=callback, **extra_args)
