"""IR-SDE sampler with the reference's API (utils/sde_utils.py:80-378 of the reference), driving the CUDA
kernels.  Drop-in for `utils.sde_utils.IRSDE` on the inference path: same constructor, attributes
(thetas, sigmas, thetas_cumsum, sigma_bars, dt, mu, model, T, sample_T, sample_scale, max_sigma) and methods
(set_mu, set_model, noise_state, score_fn, noise_fn, reverse_sde_step, reverse_posterior_step, reverse_sde,
reverse_posterior, reverse_ode).  Training-only members (forward, generate_random_states, ode_sampler,
optimal_reverse) are out of scope.

Host side: schedule tables and the per-step scalar coefficients are computed with fp32 tensor ops in the
reference's order (a few flops per step).  Device side: the state update is ONE fused kernel launch per step
(dac_sde_step) instead of ~12 pointwise kernels.

Extension over the reference (needed for parity testing, off by default): every reverse_* accepts
`noise=` - a `[T, *x.shape]` tensor (row i is consumed by the i-th executed step, i.e. t = T - i) or a
callable `noise(t) -> tensor` - so that both implementations consume identical Gaussian draws.
"""
import math
import os

import torch

from . import ops


class IRSDE:
    def __init__(self, max_sigma, T=100, sample_T=-1, schedule="cosine", eps=0.01, device=None):
        self.T = T
        self.device = device
        self.max_sigma = max_sigma / 255 if max_sigma >= 1 else max_sigma     # sde_utils.py:86
        self.sample_T = self.T if sample_T < 0 else sample_T
        self.sample_scale = self.T / self.sample_T
        self._initialize(self.max_sigma, self.sample_T, schedule, eps)

    # ------------------------------------------------------------------ schedules (sde_utils.py:90-154)
    def _initialize(self, max_sigma, T, schedule, eps):
        n = T + 1
        if schedule == "cosine":
            knots = T + 2
            grid = torch.linspace(0, knots, knots + 1, dtype=torch.float32)
            abar = torch.cos(((grid / knots) + 0.008) / (1 + 0.008) * math.pi * 0.5) ** 2
            abar = abar / abar[0]
            thetas = 1 - abar[1:-1]
        elif schedule == "linear":
            scale = 1000 / n
            thetas = torch.linspace(scale * 0.0001, scale * 0.02, n, dtype=torch.float32)
        elif schedule == "constant":
            thetas = torch.ones(n, dtype=torch.float32)
        else:
            raise NotImplementedError(f"schedule {schedule!r}")
        sigmas = torch.sqrt(max_sigma ** 2 * 2 * thetas)
        thetas_cumsum = torch.cumsum(thetas, dim=0) - thetas[0]
        self.dt = -1 / thetas_cumsum[-1] * math.log(eps)          # 0-dim CPU fp32 tensor, as in the reference
        sigma_bars = torch.sqrt(max_sigma ** 2 * (1 - torch.exp(-2 * thetas_cumsum * self.dt)))
        # host copies drive the per-step scalar coefficients; device copies keep the reference's attributes
        self._h = dict(thetas=thetas, sigmas=sigmas, thetas_cumsum=thetas_cumsum, sigma_bars=sigma_bars)
        self.thetas = thetas.to(self.device)
        self.sigmas = sigmas.to(self.device)
        self.thetas_cumsum = thetas_cumsum.to(self.device)
        self.sigma_bars = sigma_bars.to(self.device)
        self.mu = 0.
        self.model = None
        self._coef_cache = {}

    def set_mu(self, mu):
        self.mu = mu

    def set_model(self, model):
        self.model = model

    # ------------------------------------------------------------------ per-step scalars (fp32, reference order)
    def sigma_bar(self, t):
        return self.sigma_bars[t]

    def _sde_coef(self, t, half=False):
        key = ("sde", t, half)
        if key in self._coef_cache:
            return list(self._coef_cache[key])
        c = self._sde_coef_uncached(t, half)
        self._coef_cache[key] = [float(v) for v in c]
        return list(self._coef_cache[key])

    def _sde_coef_uncached(self, t, half):
        h = self._h
        s2 = h["sigmas"][t] ** 2
        if half:
            s2 = 0.5 * s2                                                     # sde_utils.py:181
        sqrt_dt = torch.tensor(math.sqrt(self.dt), dtype=torch.float32)       # python float cast to fp32 by the mul
        return [h["thetas"][t], s2, h["sigma_bars"][t], self.dt, h["sigmas"][t], sqrt_dt]

    def _posterior_coef(self, t):
        key = ("post", t)
        if key not in self._coef_cache:
            self._coef_cache[key] = [float(v) for v in self._posterior_coef_uncached(t)]
        return list(self._coef_cache[key])

    def _posterior_coef_uncached(self, t):
        h, dt = self._h, self.dt
        th, cs, cs1 = h["thetas"][t], h["thetas_cumsum"][t], h["thetas_cumsum"][t - 1]
        A, B, C_ = torch.exp(-th * dt), torch.exp(-cs * dt), torch.exp(-cs1 * dt)
        term1 = A * (1 - C_ ** 2) / (1 - B ** 2)                              # sde_utils.py:205-213
        term2 = C_ * (1 - A ** 2) / (1 - B ** 2)
        A2, B2, C2 = torch.exp(-2 * th * dt), torch.exp(-2 * cs * dt), torch.exp(-2 * cs1 * dt)
        var = (1 - A2) * (1 - C2) / (1 - B2)                                  # sde_utils.py:215-225
        logvar = torch.log(torch.clamp(var, min=1e-20 * dt))
        std = (0.5 * logvar).exp() * self.max_sigma
        return [term1, term2, std, torch.exp(cs * dt), h["sigma_bars"][t]]    # sde_utils.py:245-247

    # ------------------------------------------------------------------ single steps
    def _draw(self, x, eps):
        return torch.randn_like(x) if eps is None else eps.to(device=x.device, dtype=torch.float32)

    def _mu_like(self, x):
        mu = self.mu
        if not torch.is_tensor(mu):
            mu = torch.full_like(x, float(mu))
        return mu.to(device=x.device, dtype=torch.float32).expand_as(x).contiguous()

    def get_score_from_noise(self, noise, t):
        return -noise / self.sigma_bar(t)                                     # sde_utils.py:186-187 (API parity)

    def score_fn(self, x, t, scale=1.0, **kwargs):
        noise = self.model(x, self.mu, t * scale, **kwargs)
        return self.get_score_from_noise(noise, t)

    def noise_fn(self, x, t, scale=1.0, **kwargs):
        return self.model(x, self.mu, t * scale, **kwargs)

    def _step_from_noise(self, mode, x, net, t, eps=None, out=None):
        """x' from the network's noise prediction; one kernel launch."""
        if not x.is_cuda:
            raise ops.L.DacError("IRSDE (daclip_b200) updates CUDA tensors only (no CPU fallback)")
        with ops.L.on_device(x):
            return self._step_from_noise_on_device(mode, x, net, t, eps, out)

    def _step_from_noise_on_device(self, mode, x, net, t, eps, out):
        x = x.contiguous().float()
        net = net.float()
        out = torch.empty_like(x) if out is None else out
        if mode == "posterior":
            ops.sde_step(1, x, self._mu_like(x), net.contiguous(), self._draw(x, eps).contiguous(), out,
                         self._posterior_coef(t))
        elif mode == "sde":
            ops.sde_step(0, x, self._mu_like(x), net.contiguous(), self._draw(x, eps).contiguous(), out,
                         self._sde_coef(t))
        else:
            ops.sde_step(2, x, self._mu_like(x), net.contiguous(), None, out, self._sde_coef(t, half=True))
        return out

    def reverse_sde_step(self, x, score, t, eps=None):
        # score = -noise / sigma_bar  <=>  "noise" = score with sigma_bar = -1 (negation and /(-1) are exact)
        coef = self._sde_coef(t)
        coef[2] = -1.0
        if not x.is_cuda:
            raise ops.L.DacError("IRSDE (daclip_b200) updates CUDA tensors only (no CPU fallback)")
        with ops.L.on_device(x):
            x = x.contiguous().float()
            out = torch.empty_like(x)
            ops.sde_step(0, x, self._mu_like(x), score.contiguous().float(), self._draw(x, eps).contiguous(), out, coef)
        return out

    def reverse_ode_step(self, x, score, t):
        coef = self._sde_coef(t, half=True)
        coef[2] = -1.0
        if not x.is_cuda:
            raise ops.L.DacError("IRSDE (daclip_b200) updates CUDA tensors only (no CPU fallback)")
        with ops.L.on_device(x):
            x = x.contiguous().float()
            out = torch.empty_like(x)
            ops.sde_step(2, x, self._mu_like(x), score.contiguous().float(), None, out, coef)
        return out

    def reverse_posterior_step(self, xt, noise, t, eps=None):
        return self._step_from_noise("posterior", xt, noise, t, eps)

    def noise_state(self, tensor):
        """x_T = LQ + N(0,1) * max_sigma (sde_utils.py:374-375).  The draw uses the input's device RNG like the
        reference; the arithmetic runs on the GPU and the result returns on the input's device."""
        return self._noise_state(tensor, torch.randn_like(tensor))

    def _noise_state(self, tensor, eps):
        dev = torch.device(self.device if self.device is not None else "cuda")
        if dev.type != "cuda":
            raise ops.L.DacError("IRSDE (daclip_b200) needs a CUDA device (no CPU fallback)")
        with ops.L.on_device(dev):
            x = tensor.to(dev, torch.float32).contiguous()
            out = torch.empty_like(x)
            ops.noise_state(x, eps.to(dev, torch.float32).contiguous(), out, self.max_sigma)
        return out.to(tensor.device)

    # ------------------------------------------------------------------ loops (sde_utils.py:261-313)
    @staticmethod
    def _noise_at(noise, i, t):
        if noise is None:
            return None
        return noise(t) if callable(noise) else noise[i]

    def _save_state(self, x, t, save_dir):
        interval = self.T // 100
        if t % interval == 0:
            import torchvision.utils as tvutils
            os.makedirs(save_dir, exist_ok=True)
            x_L, x_R = x.chunk(2, dim=1)
            tvutils.save_image(torch.cat([x_L, x_R], dim=3).data, f"{save_dir}/state_{t // interval}.png",
                               normalize=False)

    def _engine_model(self, xt, kwargs):
        """The denoiser as a daclip_b200 ConditionalUNet (unwrapping DataParallel), if the fused loop applies."""
        from .unet import ConditionalUNet
        m = getattr(self.model, "module", self.model)
        ok = isinstance(m, ConditionalUNet) and torch.is_tensor(self.mu) and xt.is_cuda and xt.dim() == 4 \
            and set(kwargs) <= {"text_context", "image_context"}
        return m if ok else None

    DEVICE_LOOP = os.environ.get("DAC_DEVICE_LOOP", "1") != "0"   # one CUDA graph per sampling step, nothing else launched
    # Where the per-step Gaussian noise of the fused loop comes from when none is passed in:
    #   "philox" - generated inside the update kernel (Philox4x32-10, seeded from torch's CPU generator once per call):
    #              no noise tensor, no launch besides the step graph;
    #   "torch"  - T x torch.randn_like drawn up front in loop order, i.e. the reference's own RNG stream
    #              (sde_utils.py:227-231) - what the parity tests patch to inject the reference's draws.
    noise_source = os.environ.get("DAC_NOISE_SOURCE", "philox")

    def _reverse_fused(self, net, mode, xt, T, save_states, save_dir, noise, kwargs):
        """Hot loop: the state lives in the engine's static buffer; per step = one CUDA-graph replay of the
        denoiser plus one in-place fused update kernel.  No per-step allocation, copy or host sync."""
        with ops.L.on_device(xt):
            return self._reverse_fused_on_device(net, mode, xt, T, save_states, save_dir, noise, kwargs)

    def _reverse_fused_on_device(self, net, mode, xt, T, save_states, save_dir, noise, kwargs):
        B, _, H, W = xt.shape
        eng = net.engine(B, H, W)
        eng.set_inputs(xt, self.mu, kwargs.get("text_context"), kwargs.get("image_context"))
        code = {"sde": 0, "posterior": 1, "ode": 2}[mode]
        ts = list(reversed(range(1, T + 1)))
        eps_all, device_loop = None, self.DEVICE_LOOP and not save_states and eng.xt.numel() % 4 == 0 and len(ts) > 0
        if device_loop and noise is not None and mode != "ode":
            if isinstance(noise, (list, tuple)) and len(noise) >= len(ts) and all(torch.is_tensor(n) for n in noise[:len(ts)]):
                noise = torch.stack([n.to(eng.xt.device, torch.float32) for n in noise[:len(ts)]])
            if torch.is_tensor(noise) and noise.dim() == eng.xt.dim() + 1 and noise.shape[0] >= len(ts) \
                    and noise.shape[1:] == eng.xt.shape:
                eps_all = noise[:len(ts)].to(eng.xt.device, torch.float32).contiguous()
            else:
                device_loop = False                      # a callable / oddly shaped noise source: the per-step host loop
        elif device_loop and noise is None and mode != "ode" and self.noise_source == "torch":
            # the reference's RNG stream: T draws of torch.randn_like in loop order (sde_utils.py:227-231), made up front
            eps_all = torch.stack([torch.randn_like(eng.xt) for _ in ts])
        if device_loop:
            # The whole step - tick (time and coefficients of step s from device tables), evaluation, fused in-place
            # update with in-kernel Philox noise - is ONE CUDA graph; the host only replays it T times
            # (sde_utils.py:297-313 is a python loop with a randn_like and a dozen pointwise launches per step).
            pad = lambda c: list(c) + [0.0] * (8 - len(c))
            coefs = [pad(self._posterior_coef(t) if mode == "posterior" else self._sde_coef(t, half=(mode == "ode"))) for t in ts]
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())     # follows torch.manual_seed (CPU generator)
            eng.loop_begin([t * self.sample_scale for t in ts], coefs, eps_all, seed)
            for _ in ts:
                eng.loop_step(code)
            return eng.xt.clone()
        for i, t in enumerate(ts):
            eng.set_time(t * self.sample_scale)
            eng.replay()
            eps = None
            if mode != "ode":
                eps = self._draw(eng.xt, self._noise_at(noise, i, t)).contiguous()
            coef = self._posterior_coef(t) if mode == "posterior" else self._sde_coef(t, half=(mode == "ode"))
            ops.sde_step(code, eng.xt, eng.cond, eng.out_noise, eps, eng.xt, coef)
            if save_states:
                self._save_state(eng.xt, t, save_dir)
        return eng.xt.clone()

    def _reverse(self, mode, xt, T, save_states, save_dir, noise, kwargs):
        T = self.sample_T if T < 0 else T
        net = self._engine_model(xt, kwargs)
        if net is not None:
            return self._reverse_fused(net, mode, xt, T, save_states, save_dir, noise, kwargs)
        x = xt.clone().contiguous()
        for i, t in enumerate(reversed(range(1, T + 1))):
            net = self.model(x, self.mu, t * self.sample_scale, **kwargs)
            x = self._step_from_noise(mode, x, net, t, self._noise_at(noise, i, t))
            if save_states:
                self._save_state(x, t, save_dir)
        return x

    def reverse_sde(self, xt, T=-1, save_states=False, save_dir="sde_state", noise=None, **kwargs):
        return self._reverse("sde", xt, T, save_states, save_dir, noise, kwargs)

    def reverse_ode(self, xt, T=-1, save_states=False, save_dir="ode_state", noise=None, **kwargs):
        return self._reverse("ode", xt, T, save_states, save_dir, None, kwargs)

    def reverse_posterior(self, xt, T=-1, save_states=False, save_dir="posterior_state", noise=None, **kwargs):
        return self._reverse("posterior", xt, T, save_states, save_dir, noise, kwargs)
