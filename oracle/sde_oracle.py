"""Oracle: IR-SDE schedules and reverse-time samplers (fp32, torch, CPU).

TEST INFRASTRUCTURE - see oracle/__init__.py.  Restates
/root/reference/universal-image-restoration/utils/sde_utils.py (cited per
function as sde_utils.py:<line>).  Written from the formulas, not copied:
the reference interleaves the math with tqdm/IO; here every step is one pure
function of (x, mu, net_out, eps, t) so a CUDA kernel can be checked against
it element by element.
"""
import math

import torch


class Schedule:
    """theta / sigma / cumulative-theta / sigma-bar tables (sde_utils.py:84-154)."""

    def __init__(self, max_sigma, T=100, schedule="cosine", eps=0.01):
        # sde_utils.py:86 - values >= 1 are given in 8-bit grey levels.
        self.max_sigma = max_sigma / 255 if max_sigma >= 1 else max_sigma
        self.T = T
        n = T + 1  # index 0 .. T, index 0 never sampled
        if schedule == "cosine":
            # sde_utils.py:112-123: squared-cosine alpha-bar over T+2 knots, first
            # and last knot trimmed so that theta_0 > 0 and theta_T < 1.
            knots = T + 2
            grid = torch.linspace(0, knots, knots + 1, dtype=torch.float32)
            abar = torch.cos(((grid / knots) + 0.008) / 1.008 * math.pi * 0.5) ** 2
            abar = abar / abar[0]
            thetas = 1 - abar[1:-1]
        elif schedule == "linear":
            # sde_utils.py:101-110
            s = 1000 / n
            thetas = torch.linspace(s * 0.0001, s * 0.02, n, dtype=torch.float32)
        elif schedule == "constant":
            # sde_utils.py:93-99
            thetas = torch.ones(n, dtype=torch.float32)
        else:
            raise ValueError(f"unknown schedule {schedule!r}")
        assert thetas.numel() == n
        self.thetas = thetas
        # sde_utils.py:128-129
        self.sigmas = torch.sqrt(self.max_sigma ** 2 * 2 * thetas)
        # sde_utils.py:144: cumulative theta re-based so that Theta_0 == 0
        self.thetas_cumsum = torch.cumsum(thetas, dim=0) - thetas[0]
        # sde_utils.py:145: dt chosen so that exp(-Theta_T dt) == eps  (0-dim fp32 tensor)
        self.dt = -1 / self.thetas_cumsum[-1] * math.log(eps)
        # sde_utils.py:131-132
        self.sigma_bars = torch.sqrt(
            self.max_sigma ** 2 * (1 - torch.exp(-2 * self.thetas_cumsum * self.dt)))


def sde_step(s: Schedule, x, mu, net_out, eps, t):
    """One reverse-SDE step (sde_utils.py:44-45,177-178,183-187).

    score = -net_out / sigma_bar_t; x - (theta_t (mu - x) - sigma_t^2 score) dt
    - sigma_t * eps * sqrt(dt).
    """
    score = -net_out / s.sigma_bars[t]
    drift = (s.thetas[t] * (mu - x) - s.sigmas[t] ** 2 * score) * s.dt
    return x - drift - s.sigmas[t] * (eps * math.sqrt(s.dt))


def ode_step(s: Schedule, x, mu, net_out, t):
    """Probability-flow step (sde_utils.py:47-48,180-181): half the score term, no noise."""
    score = -net_out / s.sigma_bars[t]
    return x - (s.thetas[t] * (mu - x) - 0.5 * s.sigmas[t] ** 2 * score) * s.dt


def posterior_coeffs(s: Schedule, t):
    """Scalar coefficients of the posterior step (sde_utils.py:205-225,245-247).

    Returns fp32 0-dim tensors (term1, term2, std, expTheta, sigma_bar) such that
        x0   = (x - mu - sigma_bar * n) * expTheta + mu
        mean = term1 (x - mu) + term2 (x0 - mu) + mu
        x'   = mean + std * eps
    """
    th, cs, cs1, dt = s.thetas[t], s.thetas_cumsum[t], s.thetas_cumsum[t - 1], s.dt
    A, B, C = torch.exp(-th * dt), torch.exp(-cs * dt), torch.exp(-cs1 * dt)
    term1 = A * (1 - C ** 2) / (1 - B ** 2)
    term2 = C * (1 - A ** 2) / (1 - B ** 2)
    A2, B2, C2 = torch.exp(-2 * th * dt), torch.exp(-2 * cs * dt), torch.exp(-2 * cs1 * dt)
    var = (1 - A2) * (1 - C2) / (1 - B2)
    logvar = torch.log(torch.clamp(var, min=1e-20 * dt))
    std = (0.5 * logvar).exp() * s.max_sigma
    return term1, term2, std, torch.exp(cs * dt), s.sigma_bars[t]


def posterior_step(s: Schedule, x, mu, net_out, eps, t):
    """One posterior-sampling step (sde_utils.py:227-231)."""
    term1, term2, std, e_theta, sbar = posterior_coeffs(s, t)
    x0 = (x - mu - sbar * net_out) * e_theta + mu
    mean = term1 * (x - mu) + term2 * (x0 - mu) + mu
    return mean + std * eps


def reverse(s: Schedule, denoiser, xt, mu, mode="posterior", noise=None, T=None, time_scale=1.0, **ctx):
    """Full reverse loop (sde_utils.py:261-313).

    ``denoiser(x, mu, t, **ctx)`` predicts the noise; ``noise`` is a ``[T, ...]``
    tensor indexed as noise[T - t] (step order) so that both sides of a parity
    check consume identical Gaussian draws.  mode: 'sde' | 'posterior' | 'ode'.
    ``time_scale`` = T / sample_T of reduced-step sampling (sde_utils.py:87-89,195-202: the tables are built for
    sample_T steps and the network is queried at t * T / sample_T).
    """
    T = s.T if T is None else T
    x = xt.clone()
    for i, t in enumerate(range(T, 0, -1)):
        n = denoiser(x, mu, float(t) * time_scale, **ctx)
        if mode == "ode":
            x = ode_step(s, x, mu, n, t)
            continue
        eps = noise[i] if noise is not None else torch.randn_like(x)
        x = sde_step(s, x, mu, n, eps, t) if mode == "sde" else posterior_step(s, x, mu, n, eps, t)
    return x
