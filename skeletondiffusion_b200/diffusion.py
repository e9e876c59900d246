"""Nonisotropic Gaussian latent diffusion with the reference's API, on sm_100a kernels.

Reference: get_cov_from_corr  src/core/diffusion/utils.py:65-86
           LatentDiffusion    src/core/diffusion/base.py:64-443
           NonisotropicGaussianDiffusion  src/core/diffusion/nonisotropic.py:71-227
Buffers (all state_dict keys of the reference) are built once on the host with the reference's
fp32 expressions; sampling runs sd_sample_loop (Denoiser kernels + the fused reverse-step kernel).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import _native as nv
from .plan import Workspace, params_key

__all__ = ["get_cov_from_corr", "LatentDiffusion", "NonisotropicGaussianDiffusion"]


# ----------------------------------------------------------------------------------------------
def _is_positive_def(m: torch.Tensor) -> bool:
    assert torch.allclose(m.transpose(-1, -2), m), "Matrix must be symmetric"
    return bool((torch.linalg.eigvals(m).real > 0).all())


def get_cov_from_corr(correlation_matrix: torch.Tensor, if_sigma_n_scale=True, sigma_n_scale="spectral",
                      if_run_as_isotropic=False, diffusion_covariance_type="skeleton-diffusion", **kwargs):
    """(Sigma_N, Lambda_N, U) from a symmetric correlation/adjacency matrix (utils.py:65-86).
    One-off host-side setup (LAPACK eigh), not a kernel."""
    n = correlation_matrix.shape[0]
    dev = correlation_matrix.device
    if if_run_as_isotropic:
        eye = torch.eye(n, device=dev)
        if diffusion_covariance_type == "skeleton-diffusion":
            return torch.zeros_like(correlation_matrix), torch.ones(n, device=dev), eye
        if diffusion_covariance_type == "anisotropic":
            return eye.clone(), torch.ones(n, device=dev), eye
        return torch.zeros_like(correlation_matrix), torch.zeros(n, device=dev), eye
    sigma = correlation_matrix
    if not _is_positive_def(sigma):                                   # utils.py:19-35
        spectral = torch.linalg.eigvals(sigma).real.abs().max()
        sigma = sigma + torch.eye(n, device=dev) * (spectral + 1e-6)
        assert int((torch.linalg.eigh(sigma)[0].abs() < 0.7e-7).sum()) == 0
    lam, u = torch.linalg.eigh(sigma, UPLO="L")
    if if_sigma_n_scale:                                              # utils.py:37-62
        if sigma_n_scale == "spectral":
            s = lam.max()
        elif sigma_n_scale == "frob":
            s = lam.sum() / n
        else:
            raise AssertionError("Not implemented")
        lam, sigma = lam / s, sigma / s
        assert torch.isclose(sigma, u @ torch.diag(lam) @ u.mT, atol=1e-6).all(), "Sigma_N must be equal to U @ Lambda_N @ U.t()"
    assert (lam > 0.7e-7).all(), f"Lambda_N must be positive definite: {lam}"
    assert _is_positive_def(sigma), "Sigma_N must be positive definite"
    return sigma, lam, u


# ----------------------------------------------------------------------------------------------
def _cosine_betas(timesteps, s=0.008):
    x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


def _linear_betas(timesteps):
    scale = 1000 / timesteps
    return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)


def _exp_betas(timesteps, factor=3.0):
    return torch.clip(torch.exp(torch.linspace(-factor, 0, timesteps + 1, dtype=torch.float64)), 0, 0.999)


def _extract(a, t, x_shape):
    return a.gather(-1, t).reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


class LatentDiffusion(nn.Module):
    """Schedule buffers + the sampling / training entry points (base.py:64-443)."""

    def __init__(self, model: nn.Module, latent_size=96, diffusion_timesteps=10, diffusion_objective="pred_x0",
                 sampling_timesteps=None, diffusion_activation="identity", diffusion_conditioning=False,
                 diffusion_loss_type="mse", objective="pred_noise", beta_schedule="cosine", beta_schedule_factor=3.0,
                 ddim_sampling_eta=0., precision: str = "fp32", **kwargs):
        super().__init__()
        if diffusion_activation != "identity":
            raise NotImplementedError("diffusion_activation='tanh' is not used by the shipped configs")
        self.activation = nn.Identity()
        self.silent = True
        self.condition = diffusion_conditioning
        self.loss_type = diffusion_loss_type
        self.statistics_pred = self.statistics_obs = None
        self.model = model
        self.channels = model.channels
        self.self_condition = model.self_condition
        self.seq_length = latent_size
        self.objective = diffusion_objective          # NB: `timesteps=` / `objective=` kwargs are ignored like in the reference (base.py:90-91)
        assert self.objective in {"pred_noise", "pred_x0", "pred_v"}
        if self.objective != "pred_x0":
            raise NotImplementedError("only diffusion_objective='pred_x0' works in the reference's nonisotropic class (nonisotropic.py:123,163)")
        if beta_schedule == "linear":
            betas = _linear_betas(diffusion_timesteps)
        elif beta_schedule == "cosine":
            betas = _cosine_betas(diffusion_timesteps)
        elif beta_schedule == "exp":
            betas = _exp_betas(diffusion_timesteps, beta_schedule_factor)
        else:
            raise ValueError(f"unknown beta schedule {beta_schedule}")
        alphas = 1. - betas
        ac = torch.cumprod(alphas, dim=0)
        ac_prev = F.pad(ac[:-1], (1, 0), value=1.)
        self.num_timesteps = int(betas.shape[0])
        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else self.num_timesteps
        assert self.sampling_timesteps <= self.num_timesteps
        self.is_ddim_sampling = self.sampling_timesteps < self.num_timesteps
        if self.is_ddim_sampling:
            raise NotImplementedError("ddim_sample is broken in the reference (base.py:396 uses `times` before assignment)")
        self.ddim_sampling_eta = ddim_sampling_eta
        self.precision = precision
        for name, val in (("betas", betas), ("alphas_cumprod", ac), ("alphas_cumprod_prev", ac_prev),
                          ("sqrt_alphas_cumprod", torch.sqrt(ac))):
            self.register_buffer(name, val.to(torch.float32))
        self._noise_calls = 0
        self.use_cuda_graph = bool(kwargs.get("use_cuda_graph", False))
        self._graphs = {}

    def set_normalization_statistics(self, statistics_pred, statistics_obs):
        self.statistics_pred, self.statistics_obs = statistics_pred, statistics_obs

    # ------------------------------------------------------------------ noise
    def get_noise(self, x, *args, **kwargs):
        """White N(0, I) from the library's Philox kernel (replaces torch.randn, base.py:148-158).
        The stream is keyed by torch.initial_seed() and a per-module call counter."""
        if torch.is_tensor(x):
            shape, device = tuple(x.shape), x.device
        else:
            shape, device = tuple(x), kwargs.get("device", self.betas.device)
        out = torch.empty(shape, device=device, dtype=torch.float32)
        nv.require_cuda(out, "noise")
        seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * (self._noise_calls + 1)) & 0xFFFFFFFFFFFFFFFF
        self._noise_calls += 1
        nv.check(nv.load().sd_fill_normal(out.data_ptr(), out.numel(), seed, self._philox_offset(kwargs.get("offset", 0)),
                                          nv.stream_ptr(out.device)), "sd_fill_normal")
        return out

    @staticmethod
    def _philox_offset(element_offset: int) -> int:
        """Element offset into the conceptual global noise tensor -> Philox counter offset (one counter per 4 elements)."""
        element_offset = int(element_offset)
        if element_offset % 4:
            raise ValueError(f"noise offset must be a multiple of 4 elements, got {element_offset}")
        return element_offset // 4

    get_white_noise = get_noise
    get_start_noise = get_noise

    def fill_noise_(self, out: torch.Tensor, offset: int = 0) -> torch.Tensor:
        """White N(0, I) written into an existing contiguous fp32 CUDA tensor (same Philox stream as get_noise); `offset`: index of
        out's first element in the conceptual global tensor (a shard of rows draws what the unsharded call would draw there)."""
        nv.require_cuda(out, "noise")
        assert out.dtype == torch.float32 and out.is_contiguous()
        seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * (self._noise_calls + 1)) & 0xFFFFFFFFFFFFFFFF
        self._noise_calls += 1
        nv.check(nv.load().sd_fill_normal(out.data_ptr(), out.numel(), seed, self._philox_offset(offset), nv.stream_ptr(out.device)), "sd_fill_normal")
        return out

    # ------------------------------------------------------------------ network interface
    def feed_model(self, x, t, x_self_cond=None, x_cond=None):
        if self.condition:
            assert x_cond is not None
        return self.model(x, t, x_self_cond, x_cond, precision=self.precision)     # repeat of x_cond handled in place

    # ------------------------------------------------------------------ forward process
    def forward(self, x, *args, x_cond=None, **kwargs):
        b, c, n = x.shape
        assert n == self.seq_length, f"seq length must be {self.seq_length}"
        t = torch.randint(0, self.num_timesteps, (b,), device=x.device).long()
        return self.p_losses(x, t, *args, x_cond=x_cond, **kwargs)

    @torch.no_grad()
    def sample(self, batch_size=16, *args, **kwargs):
        return self.p_sample_loop((batch_size, self.channels, self.seq_length), *args, **kwargs)


class NonisotropicGaussianDiffusion(LatentDiffusion):
    def __init__(self, Sigma_N: torch.Tensor, Lambda_N: torch.Tensor, U: torch.Tensor,
                 diffusion_covariance_type="skeleton-diffusion", loss_reduction_type="l1", gamma_scheduler="cosine", **kwargs):
        super().__init__(**kwargs)
        Sigma_N, Lambda_N, U = Sigma_N.detach().cpu().float(), Lambda_N.detach().cpu().float(), U.detach().cpu().float()
        # torch.linalg.eigh returns U in column-major strides: every derived table would inherit them, and the
        # kernels take raw pointers, so all buffers are registered row-major contiguous
        reg = lambda name, val: self.register_buffer(name, val.to(torch.float32).contiguous())
        reg("Lambda_N", Lambda_N)
        reg("Sigma_N", Sigma_N)
        reg("U", U)
        reg("U_transposed", U.t())
        betas, ac, ac_prev = self.betas, self.alphas_cumprod, self.alphas_cumprod_prev
        alphas = 1. - betas
        n = Lambda_N.shape[0]
        # per-eigenmode variance schedules (nonisotropic.py:36-68), fp32 like the reference
        if diffusion_covariance_type == "anisotropic":
            lam_t = (1 - alphas.unsqueeze(-1)) * Lambda_N
            lam_bar = (1 - ac.unsqueeze(-1)) * Lambda_N
            lam_bar_prev = (1 - ac_prev.unsqueeze(-1)) * Lambda_N
        elif diffusion_covariance_type == "skeleton-diffusion":
            if gamma_scheduler == "cosine":
                gammas = 1 - alphas
            elif gamma_scheduler == "mono_decrease":
                gammas = 1 - torch.arange(0, self.num_timesteps) / self.num_timesteps
            else:
                raise AssertionError("Not implemented")
            lam_i = Lambda_N - 1
            g_bar = (1 - alphas) * gammas
            g_tilde = ac * torch.cumsum(g_bar / ac, dim=-1)
            lam_t = lam_i.unsqueeze(0) * g_bar.unsqueeze(-1) + (1 - alphas).unsqueeze(-1)
            lam_bar = lam_i.unsqueeze(0) * g_tilde.unsqueeze(-1) + (1 - ac.unsqueeze(-1))
            lam_bar_prev = torch.cat([torch.zeros(n).unsqueeze(0), lam_bar[:-1]], dim=0)
        else:
            raise NotImplementedError("diffusion_covariance_type='isotropic' crashes in the reference ctor (nonisotropic.py:108); "
                                      "use get_cov_from_corr(if_run_as_isotropic=True) with the default covariance type")
        ut = self.U_transposed
        inv_sqrt = 1 / torch.sqrt(lam_bar)
        reg("inv_sqrt_Lambda_bar_mmUt", inv_sqrt.unsqueeze(-1) * ut.unsqueeze(0))
        reg("inv_sqrt_Lambda_bar_sqrt_alphas_cumprod_mmUt", (inv_sqrt * self.sqrt_alphas_cumprod.unsqueeze(-1)).unsqueeze(-1) * ut.unsqueeze(0))
        reg("Umm_sqrt_Lambda_bar_t", U.unsqueeze(0) * torch.sqrt(lam_bar).unsqueeze(-2))
        reg("Umm_sqrt_Lambda_bar_t_sqrt_recip_alphas_cumprod", U.unsqueeze(0) * torch.sqrt(lam_bar / ac.unsqueeze(-1)).unsqueeze(-2))
        lam_post = lam_t * lam_bar_prev * (1 / lam_bar)
        reg("Lambda_posterior", lam_post)
        reg("Lambda_posterior_log_variance_clipped", torch.log(lam_post.clamp(min=1e-20)))
        diag = lambda v: torch.stack([torch.diag(d) for d in v], dim=0)
        reg("posterior_mean_coef1_x0", torch.sqrt(ac_prev)[:, None, None] * (U.unsqueeze(0) @ diag((1 / lam_bar) * lam_t) @ ut.unsqueeze(0)))
        reg("posterior_mean_coef2_xt", torch.sqrt(alphas)[:, None, None] * (U.unsqueeze(0) @ diag((1 / lam_bar) * lam_bar_prev) @ ut.unsqueeze(0)))
        self.loss_reduction_type = loss_reduction_type
        reg("mahalanobis_S_sqrt_recip", torch.sqrt(1. / lam_bar).unsqueeze(-1) * ut.unsqueeze(0))
        reg("loss_weight", ac.clone())                      # objective pred_x0 (nonisotropic.py:120-121)
        self._plan = None

    def check_eigh(self):
        return torch.isclose(self.U @ torch.diag(self.Lambda_N) @ self.U_transposed, self.Sigma_N)

    # ------------------------------------------------------------------ native plan
    def _diffusion_plan(self):
        tabs = [self.posterior_mean_coef1_x0, self.posterior_mean_coef2_xt, self.Lambda_posterior_log_variance_clipped, self.U]
        key = params_key(tabs)
        if self._plan is None or self._plan["key"] != key:
            nv.require_cuda(self.U, "diffusion buffers (call .to('cuda'))")
            c1h = self.posterior_mean_coef1_x0.detach().cpu().contiguous()
            c2h = self.posterior_mean_coef2_xt.detach().cpu().contiguous()
            logv = self.Lambda_posterior_log_variance_clipped.detach().cpu()
            sh = (self.U.detach().cpu().unsqueeze(0) * (0.5 * logv).exp().unsqueeze(-2)).contiguous()     # U diag(sigma_t)
            dev = self.U.device
            c1, c2, s = c1h.to(dev), c2h.to(dev), sh.to(dev)
            handle = C.c_void_p()
            nv.check(nv.load().sd_diffusion_create(c1.shape[1], self.seq_length, self.num_timesteps, c1.data_ptr(), c2.data_ptr(),
                                                   s.data_ptr(), c1h.data_ptr(), c2h.data_ptr(), sh.data_ptr(), C.byref(handle)),
                     "sd_diffusion_create")
            if self._plan is not None:
                nv.load().sd_diffusion_destroy(self._plan["handle"])
            self._plan = dict(key=key, handle=handle, c1=c1, c2=c2, s=s)
        return self._plan

    def __del__(self):
        try:
            if getattr(self, "_plan", None):
                nv.load().sd_diffusion_destroy(self._plan["handle"])
                self._plan = None
        except Exception:
            pass

    # ------------------------------------------------------------------ forward process / loss
    def q_sample(self, x_start, t, noise=None):
        noise = noise if noise is not None else self.get_white_noise(x_start)
        nv.require_cuda(x_start, "x_start")
        x_start, noise = x_start.float().contiguous(), noise.float().contiguous()
        b, n, d = x_start.shape
        out = torch.empty_like(x_start)
        t32 = t.to(x_start.device, torch.int32).contiguous()
        sqrt_ac, m = self.sqrt_alphas_cumprod.contiguous(), self.Umm_sqrt_Lambda_bar_t.contiguous()
        nv.check(nv.load().sd_q_sample(x_start.data_ptr(), noise.data_ptr(), t32.data_ptr(), sqrt_ac.data_ptr(), m.data_ptr(),
                                       out.data_ptr(), b, n, d, nv.stream_ptr(x_start.device)), "sd_q_sample")
        return out

    def p_losses(self, x_start, t, noise=None, x_cond=None, n_train_samples=1):
        """Training objective (base.py:262-300): returns (loss [B * k], loss_weight[t] [B], model_out).
        Without autograd (torch.no_grad(), or a frozen Denoiser) the values come from the inference kernels alone.  With
        autograd, k = 1 runs the differentiable Denoiser (training.py); k > 1 (TrainerDiffusion's best-of-k relaxation,
        trainer.py:224-234) computes all B * k values with the inference kernels and differentiates only the rows that receive
        a loss gradient (training.SparseRowLoss)."""
        from . import training
        b = x_start.shape[0]
        k = int(n_train_samples)
        if k > 1:
            x_start = x_start.repeat_interleave(k, dim=0)
            t = t.repeat_interleave(k, dim=0)
        noise = noise if noise is not None else self.get_white_noise(x_start)
        x = self.q_sample(x_start=x_start, t=t, noise=noise)
        x_start = x_start.float().contiguous()
        if self.loss_reduction_type != "l1":
            raise NotImplementedError("loss_reduction_type 'mse'")
        params = list(self.model.parameters())
        with_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if with_grad and k == 1:
            loss, model_out = training.diffusion_loss_train(self, x, x_start, t, x_cond)
        else:
            with torch.no_grad():
                model_out = self.feed_model(x, t, x_cond=x_cond)
                _, n, d = x_start.shape
                loss = torch.empty(x_start.shape[0], device=x_start.device, dtype=torch.float32)
                t32 = t.to(x_start.device, torch.int32).contiguous()
                s_tab = self.mahalanobis_S_sqrt_recip.contiguous()
                nv.check(nv.load().sd_mahalanobis_loss(model_out.data_ptr(), x_start.data_ptr(), t32.data_ptr(), s_tab.data_ptr(),
                                                       loss.data_ptr(), x_start.shape[0], n, d, nv.stream_ptr(x_start.device)),
                         "sd_mahalanobis_loss")
            if with_grad:
                xc = None if x_cond is None else x_cond.float().contiguous()
                rep = 1 if xc is None else x_start.shape[0] // xc.shape[0]
                loss = training.SparseRowLoss.apply(loss, self, x, x_start, t, xc, rep, *params)
        return loss, _extract(self.loss_weight, t.view(b, -1)[:, 0], loss.shape[0:1]), model_out

    # ------------------------------------------------------------------ reverse process
    def _reverse_step(self, x, x0, eps, t: int, clip_denoised=True, want_mean=False):
        plan = self._diffusion_plan()
        out = torch.empty_like(x)
        mean = torch.empty_like(x) if want_mean else None
        ev = nv.view_of(eps) if eps is not None else None
        nv.check(nv.load().sd_reverse_step(plan["handle"], x.data_ptr(), x0.data_ptr(), C.byref(ev) if ev is not None else None,
                                           out.data_ptr(), nv.dptr(mean), t, x.shape[0], 1 if clip_denoised else 0,
                                           nv.stream_ptr(x.device)), "sd_reverse_step")
        return out, mean

    # ------------------------------------------------------------------ whole loop as one CUDA graph
    def _graph_sample(self, dplan, mplan, img, x_cond, rep, sampling_noise, mean_t, B, clip_denoised, ws, prec):
        """Replays the T-step loop (~50 launches per step) as ONE captured CUDA graph.  Graphs are cached per
        (batch, conditioning geometry, precision, plan identity); inputs are copied into the graph's static buffers."""
        device = img.device
        key = (B, None if x_cond is None else tuple(x_cond.shape), rep, prec, bool(clip_denoised), mean_t is not None,
               dplan["handle"].value, mplan.handle.value, ws.data_ptr())
        entry = self._graphs.get(key)
        lib = nv.load()
        if entry is None:
            st = dict(x=torch.empty_like(img), cond=None if x_cond is None else torch.empty_like(x_cond),
                      noise=None if sampling_noise is None else torch.empty_like(sampling_noise),
                      means=None if mean_t is None else torch.empty_like(mean_t))

            def run():
                cv = nv.view_of(st["cond"], rep) if st["cond"] is not None else None
                nv.check(lib.sd_sample_loop(dplan["handle"], mplan.handle, st["x"].data_ptr(), C.byref(cv) if cv is not None else None,
                                            nv.dptr(st["noise"]), nv.dptr(st["means"]), B, 1 if clip_denoised else 0, ws.data_ptr(),
                                            prec, nv.stream_ptr(device)), "sd_sample_loop")

            for name, src in (("x", img), ("cond", x_cond), ("noise", sampling_noise)):
                if st[name] is not None:
                    st[name].copy_(src)
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                run()                                   # warm-up outside capture (function attributes, tensor maps)
            torch.cuda.current_stream(device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                run()
            # the entry keeps what the capture baked in alive (plans, workspace: their addresses cannot be reused while it
            # lives) and entries captured for plans that have been replaced since are dropped
            live = (dplan["handle"].value, mplan.handle.value)
            for k_old in [k for k, e in self._graphs.items() if e["plans"] != live]:
                del self._graphs[k_old]
            entry = self._graphs[key] = dict(graph=graph, st=st, plans=live, refs=(dplan, mplan, ws))
        st = entry["st"]
        st["x"].copy_(img)
        if x_cond is not None:
            st["cond"].copy_(x_cond)
        if sampling_noise is not None:
            st["noise"].copy_(sampling_noise)
        entry["graph"].replay()
        if mean_t is not None:
            mean_t.copy_(st["means"])
        return st["x"].clone()

    @torch.no_grad()
    def p_sample(self, x, t: int, x_self_cond=None, clip_denoised=True, sampling_noise=None, *args, x_cond=None,
                 if_interpolate=False, noise2interpolate=None, interpolation_kwargs: Dict = None, **kwargs):
        """One reverse step (base.py:324-341): returns (x_{t-1}, clamp(x0), noise, posterior mean)."""
        x = x.float().contiguous()
        b = x.shape[0]
        times = torch.full((b,), t, device=x.device, dtype=torch.long)
        x0 = self.feed_model(x, times, x_self_cond, x_cond)
        if sampling_noise is not None and t > 0:
            noise = sampling_noise[:, sampling_noise.shape[1] - t]
        else:
            noise = self.get_white_noise(x) if t > 0 else None
        if if_interpolate and t > 0:
            noise2 = noise2interpolate[:, sampling_noise.shape[1] - t]
            zeros = torch.zeros_like(x)
            n1, _ = self._reverse_step(zeros, zeros, noise, t, clip_denoised=False)         # U (sigma_t * eps)
            n2, _ = self._reverse_step(zeros, zeros, noise2, t, clip_denoised=False)
            mean, _ = self._reverse_step(x, x0, None, t, clip_denoised)
            img = mean + interpolation_kwargs["interpolate_funct"](n1, n2)                  # nonisotropic.py:218-227
        else:
            img, mean = self._reverse_step(x, x0, noise, t, clip_denoised, want_mean=True)
        x_start = x0.clamp(-1., 1.) if clip_denoised else x0
        return img, x_start, (noise if noise is not None else 0.), mean

    @torch.no_grad()
    def p_sample_loop(self, shape, x_cond=None, start_noise=None, sampling_noise=None, return_sampling_noise=False,
                      return_timages=False, clip_denoised=True, if_interpolate=False, noise_row_offset: int = 0, **kwargs):
        """Full reverse process (base.py:343-390).  The common case runs as ONE native call
        (sd_sample_loop: T x (Denoiser kernels + fused step kernel), CUDA-graph capturable)."""
        device = self.betas.device
        nv.require_cuda(self.betas, "diffusion buffers (call .to('cuda'))")
        B, N, D = shape
        T = self.num_timesteps
        if start_noise is not None:
            assert tuple(start_noise.shape) == tuple(shape), f"Shape mismatch: {start_noise.shape} != {shape}"
            img = start_noise.to(device, torch.float32)
        else:
            # noise_row_offset: global index of this call's first row (a rank's window shard x samples): the Philox stream
            # is indexed by the global row, so the drawn noise does not depend on how the windows are sharded over ranks
            img = self.get_start_noise(tuple(shape), device=device, offset=int(noise_row_offset) * N * D)
        noise0 = img.clone()
        if sampling_noise is not None:
            assert tuple(sampling_noise.shape) == (B, T - 1, N, D), f"Shape mismatch: {tuple(sampling_noise.shape)}"
            sampling_noise = sampling_noise.to(device, torch.float32).contiguous()
        elif T > 1:
            sampling_noise = self.get_white_noise((B, T - 1, N, D), device=device, offset=int(noise_row_offset) * (T - 1) * N * D)
        if self.condition:
            assert x_cond is not None
        if return_timages or if_interpolate:                      # rarely used variants: per-step calls
            imgs, means = [], []
            for t in reversed(range(T)):
                img, _, _, mean = self.p_sample(img, t, None, clip_denoised, sampling_noise, x_cond=x_cond,
                                                if_interpolate=if_interpolate, **kwargs)
                if t != 0:
                    imgs.append(img)
                    means.append(mean)
            noise_t = sampling_noise
            mean_t = torch.stack(means, 1) if means else None
            imgs = torch.stack(imgs, 1) if imgs else None
        else:
            dplan, mplan = self._diffusion_plan(), self.model.plan(min_time_rows=T)
            img = img.contiguous().clone() if start_noise is not None else img
            rep = 1
            if x_cond is not None:
                x_cond = x_cond.to(device, torch.float32)
                if x_cond.stride(-1) != 1:
                    x_cond = x_cond.contiguous()
                if x_cond.shape[0] != B:
                    assert B % x_cond.shape[0] == 0
                    rep = B // x_cond.shape[0]
            mean_t = torch.empty(B, T - 1, N, D, device=device, dtype=torch.float32) if return_sampling_noise else None
            lib = nv.load()
            prec = nv.PRECISIONS[self.precision]
            ws = Workspace.get(device, lib.sd_sample_workspace_bytes(dplan["handle"], mplan.handle, B, prec), "sample")
            if self.use_cuda_graph:
                img = self._graph_sample(dplan, mplan, img, x_cond, rep, sampling_noise, mean_t, B, clip_denoised, ws, prec)
            else:
                cv = nv.view_of(x_cond, rep) if x_cond is not None else None
                nv.check(lib.sd_sample_loop(dplan["handle"], mplan.handle, img.data_ptr(), C.byref(cv) if cv is not None else None,
                                            nv.dptr(sampling_noise), nv.dptr(mean_t), B, 1 if clip_denoised else 0, ws.data_ptr(),
                                            prec, nv.stream_ptr(device)), "sd_sample_loop")
            noise_t, imgs = sampling_noise, None
        noise = noise0
        if return_sampling_noise:
            noise = (noise0, noise_t, imgs) if return_timages else (noise0, noise_t, mean_t)
        elif return_timages:
            noise = (noise0, imgs)
        return img, noise
