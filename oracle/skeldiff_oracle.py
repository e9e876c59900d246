"""CPU oracle for the SkeletonDiffusion nonisotropic sampling path.

TEST INFRASTRUCTURE ONLY.  This file is a plain PyTorch fp32 (CPU) restatement of the
reference algorithm, written as pure functions over a flat ``state_dict``.  It is imported only
by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs.  The product (``skeletondiffusion_b200``) never imports it and has no CPU fallback.

Parity status: PINNED.  Every function below is checked (tests/test_oracle_golden.py) against
golden vectors produced by running the reference's own classes from /root/reference in the build
container (generator: tests/golden/make_golden.py, fixtures: tests/golden/*.npz).  The only
third-party arithmetic that is not under /root/reference is
``denoising-diffusion-pytorch==1.9.4::SinusoidalPosEmb`` (README.md:151, used at
src/core/network/nn/generator.py:47); it is restated from the public definition and is
"parity unpinned" at that one boundary.

All citations are file:line under /root/reference/.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------
# covariance from a correlation / adjacency matrix        src/core/diffusion/utils.py:19-86
# --------------------------------------------------------------------------------------------
def cov_from_corr(corr: Tensor, if_sigma_n_scale: bool = True, sigma_n_scale: str = "spectral",
                  if_run_as_isotropic: bool = False,
                  diffusion_covariance_type: str = "skeleton-diffusion") -> Tuple[Tensor, Tensor, Tensor]:
    n = corr.shape[0]
    if if_run_as_isotropic:                                            # utils.py:68-80
        eye = torch.eye(n)
        if diffusion_covariance_type == "skeleton-diffusion":
            return torch.zeros_like(corr), torch.ones(n), eye
        if diffusion_covariance_type == "anisotropic":
            return eye.clone(), torch.ones(n), eye
        return torch.zeros_like(corr), torch.zeros(n), eye
    ev = torch.linalg.eigvals(corr)                                    # utils.py:21
    if not bool((ev.real > 0).all()):                                  # utils.py:23-30
        corr = corr + torch.eye(n) * (ev.real.abs().max() + 1e-6)
    lam, u = torch.linalg.eigh(corr, UPLO="L")                         # utils.py:83
    if if_sigma_n_scale:                                               # utils.py:43-54
        s = lam.max() if sigma_n_scale == "spectral" else lam.sum() / n
        lam = lam / s
        corr = corr / s
    return corr, lam, u


# --------------------------------------------------------------------------------------------
# schedules and per-step tables            base.py:45-55,112-134; nonisotropic.py:36-121
# --------------------------------------------------------------------------------------------
def cosine_betas(timesteps: int, s: float = 0.008) -> Tensor:
    x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)          # base.py:50-55
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    return torch.clip(1 - ac[1:] / ac[:-1], 0, 0.999)


def diffusion_tables(lambda_n: Tensor, u: Tensor, timesteps: int = 10) -> Dict[str, Tensor]:
    """fp32 tables of the 'skeleton-diffusion' covariance type with the cosine gamma scheduler."""
    betas64 = cosine_betas(timesteps)
    ac64 = torch.cumprod(1.0 - betas64, 0)
    betas = betas64.float()                                            # base.py:127-132
    ac = ac64.float()
    ac_prev = F.pad(ac64[:-1], (1, 0), value=1.0).float()
    alphas = 1.0 - betas
    lam_i = lambda_n.float() - 1                                       # nonisotropic.py:61
    g_bar = (1 - alphas) * (1 - alphas)                                # :54,:62
    g_tilde = ac * torch.cumsum(g_bar / ac, dim=-1)                    # :63
    lam_t = lam_i[None] * g_bar[:, None] + (1 - alphas)[:, None]       # :64
    lam_bar = lam_i[None] * g_tilde[:, None] + (1 - ac[:, None])       # :65
    lam_bar_prev = torch.cat([torch.zeros(1, lam_bar.shape[1]), lam_bar[:-1]], 0)   # :66
    lam_post = lam_t * lam_bar_prev * (1 / lam_bar)                    # :103
    ut = u.t()
    diag = lambda v: torch.stack([torch.diag(d) for d in v], 0)
    c1 = torch.sqrt(ac_prev)[:, None, None] * (u[None] @ diag((1 / lam_bar) * lam_t) @ ut[None])        # :108
    c2 = torch.sqrt(alphas)[:, None, None] * (u[None] @ diag((1 / lam_bar) * lam_bar_prev) @ ut[None])  # :109
    return dict(betas=betas, alphas_cumprod=ac, alphas_cumprod_prev=ac_prev,
                sqrt_alphas_cumprod=torch.sqrt(ac64).float(),
                Lambda_t=lam_t, Lambda_bar_t=lam_bar, Lambda_posterior=lam_post,
                Lambda_posterior_log_variance_clipped=torch.log(lam_post.clamp(min=1e-20)),
                posterior_mean_coef1_x0=c1, posterior_mean_coef2_xt=c2,
                Umm_sqrt_Lambda_bar_t=u[None] * torch.sqrt(lam_bar)[:, None, :],          # :98
                mahalanobis_S_sqrt_recip=torch.sqrt(1.0 / lam_bar)[:, :, None] * ut[None],  # :115-116
                loss_weight=ac)                                        # :120-121 (pred_x0)


# --------------------------------------------------------------------------------------------
# StaticGraphLinear                                     layers/graph_structural.py:30-43
# --------------------------------------------------------------------------------------------
def graph_linear(sd: SD, p: str, x: Tensor, node_types: Optional[Tensor], learn_influence: bool) -> Tensor:
    g = sd[p + "G"]
    if learn_influence:
        g = F.normalize(g, p=1.0, dim=1)                               # :32
    w = sd[p + "weight"]
    if node_types is not None:
        y = torch.einsum("ndo,bnd->bno", w[node_types].transpose(-2, -1), x)   # :36-37, :7-8
    else:
        y = torch.matmul(x, w.t())
    if (p + "bias") in sd:
        b = sd[p + "bias"]
        y = y + (b[node_types] if node_types is not None else b)       # :38-40
    return g.matmul(y)                                                 # :41


# --------------------------------------------------------------------------------------------
# Denoiser                           nn/generator.py:86-107, layers/attention.py:30-136
# --------------------------------------------------------------------------------------------
def sinusoidal_embedding(t: Tensor, dim: int, theta: float = 10000.0) -> Tensor:
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(theta) / (half - 1)))
    ang = t.float()[:, None] * freq[None]
    return torch.cat([ang.sin(), ang.cos()], -1)


def _resnet_block(sd: SD, p: str, x: Tensor, temb: Tensor, nt, li, has_res_linear: bool) -> Tensor:
    ss = F.linear(torch.tanh(temb), sd[p + "mlp.1.weight"], sd[p + "mlp.1.bias"])     # attention.py:81-84,93-94
    scale, shift = ss[:, None, :].chunk(2, dim=-1)                                     # :95-96
    h = graph_linear(sd, p + "block1.proj.", x, nt, li)
    h = torch.tanh(h * (scale + 1) + shift)                                            # :70-74
    h = torch.tanh(graph_linear(sd, p + "block2.proj.", h, nt, li))                    # :100
    res = graph_linear(sd, p + "res_linear.", x, nt, li) if has_res_linear else x     # :88,:102
    return h + res


def _node_attention(sd: SD, p: str, x: Tensor, heads: int, dim_head: int, nt, li) -> Tensor:
    b, n, c = x.shape
    y = F.normalize(x, dim=-1) * sd[p + "norm.g"] * (c ** 0.5)                         # attention.py:36
    qkv = graph_linear(sd, p + "fn.to_qkv.", y, nt, li)                                # :124
    q, k, v = [t.reshape(b, n, heads, dim_head).permute(0, 2, 1, 3) for t in qkv.chunk(3, -1)]   # b h n c
    sim = torch.einsum("bhnc,bhjc->bhnj", q * dim_head ** -0.5, k)                     # :128-129
    attn = sim.softmax(-1)                                                             # :130
    out = torch.einsum("bhnj,bhjd->bhnd", attn, v)                                     # :133
    out = out.permute(0, 2, 1, 3).reshape(b, n, heads * dim_head)                      # :135
    return graph_linear(sd, p + "fn.to_out.", out, nt, li) + x                         # :136, :16-17


def denoiser_forward(sd: SD, cfg: dict, x: Tensor, time: Tensor, x_cond: Optional[Tensor] = None,
                     prefix: str = "") -> Tensor:
    """cfg keys: dim, cond_dim, depth, attn_heads, attn_dim_head, node_types (LongTensor|None),
    learn_influence, sinusoidal_pos_emb_theta (default 1e4)."""
    nt, li = cfg.get("node_types"), cfg.get("learn_influence", False)
    heads, dh = cfg.get("attn_heads", 4), cfg.get("attn_dim_head", 32)
    c = cfg["dim"] + cfg.get("cond_dim", 0)
    p = prefix
    if x_cond is not None:
        x = torch.cat([x_cond, x], -1)                                                 # generator.py:91-92
    x = graph_linear(sd, p + "init_lin.", x, nt, li)                                   # :94
    r = x.clone()
    temb = sinusoidal_embedding(time, c, cfg.get("sinusoidal_pos_emb_theta", 10000.0))
    temb = F.linear(temb, sd[p + "time_mlp.1.weight"], sd[p + "time_mlp.1.bias"])
    temb = F.linear(F.gelu(temb), sd[p + "time_mlp.3.weight"], sd[p + "time_mlp.3.bias"])   # :50-55,:97
    n_pairs = 2 * cfg.get("depth", 1)
    for i in range(n_pairs):                                                           # :100-102
        x = _resnet_block(sd, f"{p}layers.{i}.0.", x, temb, nt, li, False)
        if i != n_pairs - 1:                                                           # :69-77
            x = _node_attention(sd, f"{p}layers.{i}.1.fn.", x, heads, dh, nt, li)
    x = torch.cat([x, r], -1)                                                          # :104
    x = _resnet_block(sd, p + "final_res_block.", x, temb, nt, li, True)               # :106
    return graph_linear(sd, p + "final_glin.", x, nt, li)                              # :107


# --------------------------------------------------------------------------------------------
# reverse process                      base.py:314-390; nonisotropic.py:196-210
# --------------------------------------------------------------------------------------------
def reverse_step(tab: Dict[str, Tensor], u: Tensor, x_t: Tensor, x0: Tensor, t: int,
                 noise: Optional[Tensor], clip_denoised: bool = True) -> Tuple[Tensor, Tensor]:
    """One reverse step given the model output x0.  Returns (x_{t-1}, posterior mean)."""
    if clip_denoised:
        x0 = x0.clamp(-1.0, 1.0)                                                       # base.py:318-319
    mean = tab["posterior_mean_coef1_x0"][t] @ x0 + tab["posterior_mean_coef2_xt"][t] @ x_t   # nonisotropic.py:196-200
    if noise is None or t == 0:                                                        # base.py:333
        return mean, mean
    logvar = tab["Lambda_posterior_log_variance_clipped"][t][:, None]                  # nonisotropic.py:205
    return mean + u @ ((0.5 * logvar).exp() * noise), mean                             # :208-210


def sample(sd: SD, cfg: dict, tab: Dict[str, Tensor], u: Tensor, x_cond: Optional[Tensor],
           start_noise: Tensor, sampling_noise: Tensor, prefix: str = "model.",
           return_means: bool = False):
    """p_sample_loop with injected noise (base.py:343-390).  sampling_noise: [B, T-1, N, D]."""
    timesteps = tab["betas"].shape[0]
    img = start_noise
    means = []
    for t in reversed(range(timesteps)):
        tt = torch.full((img.shape[0],), t, dtype=torch.long)                          # base.py:327
        x0 = denoiser_forward(sd, cfg, img, tt, x_cond, prefix)
        noise = sampling_noise[:, sampling_noise.shape[1] - t] if t > 0 else None      # :330-333
        img, mean = reverse_step(tab, u, img, x0, t, noise)
        if t != 0:
            means.append(mean)
    if return_means:
        return img, torch.stack(means, 1)
    return img


def q_sample(tab: Dict[str, Tensor], x_start: Tensor, t: Tensor, noise: Tensor) -> Tensor:
    """nonisotropic.py:152-159."""
    a = tab["sqrt_alphas_cumprod"][t][:, None, None]
    return a * x_start + tab["Umm_sqrt_Lambda_bar_t"][t] @ noise


def p_losses(sd: SD, cfg: dict, tab: Dict[str, Tensor], x_start: Tensor, t: Tensor, noise: Tensor,
             x_cond: Optional[Tensor] = None, prefix: str = "model."):
    """Training loss, objective pred_x0, loss_reduction 'l1' (base.py:262-300; nonisotropic.py:180-190)."""
    x = q_sample(tab, x_start, t, noise)
    out = denoiser_forward(sd, cfg, x, t, x_cond, prefix)
    loss = (tab["mahalanobis_S_sqrt_recip"][t] @ (out - x_start)).abs()
    return loss.mean(dim=(1, 2)), tab["loss_weight"][t], out


# --------------------------------------------------------------------------------------------
# graph GRU and the autoencoder     layers/recurrent.py:321-395; nn/encoder.py, nn/decoder.py
# --------------------------------------------------------------------------------------------
def gru_cell(sd: SD, p: str, x: Tensor, h: Tensor, gx: Tensor, nt: Optional[Tensor]) -> Tensor:
    w_ih, w_hh = sd[p + "weight_ih"], sd[p + "weight_hh"]
    b_ih, b_hh = sd[p + "bias_ih"], sd[p + "bias_hh"]
    if nt is not None:                                                                 # recurrent.py:333-349
        xr = torch.einsum("ndo,bnd->bno", w_ih[nt].transpose(-2, -1), x) + b_ih[nt]
        hr = torch.einsum("ndo,bnd->bno", w_hh[nt].transpose(-2, -1), h) + b_hh[nt]
    else:
        xr = x @ w_ih.t() + b_ih
        hr = h @ w_hh.t() + b_hh
    xr, hr = gx @ xr, gx @ hr
    i_r, i_z, i_n = xr.chunk(3, 2)
    h_r, h_z, h_n = hr.chunk(3, 2)
    r = torch.sigmoid(i_r + h_r)                                                       # :354-356
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return n - n * z + z * h                                                           # :358 (clock mask == 1)


def _next_gx(gx: Tensor, g_add) -> Tensor:
    return F.normalize(gx + g_add, p=1.0, dim=1)                                       # recurrent.py:361-363


def encode(sd: SD, cfg: dict, obs: Tensor, prefix: str = "encoder.") -> Tensor:
    """AutoEncoder.get_past_embedding (autoencoder.py:51-55): [W,T,N,3] -> [W,N,latent]."""
    nt = cfg.get("node_types")
    p = prefix
    h0 = graph_linear(sd, p + "initial_hidden1.", obs[:, 0], nt, True)                 # encoder.py:66
    seq = obs
    for layer in range(cfg.get("enc_num_layers", 1)):                                  # recurrent.py:384-394
        lp = f"{p}rnn.layers.{layer}."
        gx = F.normalize(sd[lp + "G"], p=1.0, dim=1)
        h = h0
        outs = []
        for t in range(seq.shape[1]):
            h = gru_cell(sd, lp, seq[:, t], h, gx, nt)
            gx = _next_gx(gx, 0.0)
            outs.append(h)
        seq = torch.stack(outs, 1)
    z = torch.tanh(graph_linear(sd, p + "fc.", seq[:, -1], nt, True))                  # encoder.py:81
    return torch.tanh(z)                                                               # autoencoder.py:54


def decode(sd: SD, cfg: dict, x_last2: Tensor, latent: Tensor, ph: int, prefix: str = "decoder.") -> Tensor:
    """AutoEncoder.decode (autoencoder.py:66-73; decoder.py:61-104): x_last2 [B,2,N,3] -> [B,ph,N,3]."""
    nt = cfg.get("node_types")
    p = prefix
    h = graph_linear(sd, p + "initial_hidden_h.", torch.cat([x_last2[:, -2], latent], -1), nt, True)   # decoder.py:65,73
    rec_in = torch.cat([x_last2[:, -1], latent], -1)                                   # :81
    lp = p + "rnn.layers.0."
    gx = F.normalize(sd[lp + "G"], p=1.0, dim=1)
    out = []
    for _ in range(ph):                                                                # :91-100
        h = gru_cell(sd, lp, rec_in, h, gx, nt)
        gx = _next_gx(gx, sd[lp + "G_add"])
        out.append(torch.tanh(graph_linear(sd, p + "fc.", h, nt, True)))
    return torch.stack(out, 1)


def get_prediction(ae_sd: SD, diff_sd: SD, cfg: dict, tab: Dict[str, Tensor], u: Tensor, obs: Tensor,
                   num_samples: int, pred_length: int, start_noise: Tensor, sampling_noise: Tensor) -> Tensor:
    """src/eval_prepare_model.py:89-121 with injected noise: obs [W,T,N,3] -> [W,S,ph,N,3]."""
    w = obs.shape[0]
    z_past = encode(ae_sd, cfg, obs)
    zc = z_past.repeat_interleave(num_samples, 0)
    lat = sample(diff_sd, cfg, tab, u, zc, start_noise, sampling_noise)
    pred = decode(ae_sd, cfg, obs[:, -2:].repeat_interleave(num_samples, 0), lat, pred_length)
    return pred.view(w, num_samples, pred_length, obs.shape[2], obs.shape[3])


# --------------------------------------------------------------------------------------------
# metrics                                              src/metrics/multimodal.py:15-73
# --------------------------------------------------------------------------------------------
def apd(pred: Tensor) -> Tensor:
    w, s = pred.shape[:2]
    arr = pred.reshape(w, s, -1)
    dist = torch.cdist(arr, arr)
    iu = torch.triu_indices(s, s, offset=1)
    return dist[:, iu[0], iu[1]].mean(-1)


def ade(target: Tensor, pred: Tensor) -> Tensor:
    w, s, t = pred.shape[:3]
    d = torch.linalg.norm(pred.reshape(w, s, t, -1) - target.reshape(w, 1, t, -1), dim=-1).mean(-1)
    return d.min(-1).values


def fde(target: Tensor, pred: Tensor) -> Tensor:
    w, s, t = pred.shape[:3]
    d = torch.linalg.norm(pred.reshape(w, s, t, -1) - target.reshape(w, 1, t, -1), dim=-1)[..., -1]
    return d.min(-1).values


def mmade(pred: Tensor, mm_gt) -> Tensor:
    """src/metrics/multimodal.py:105-120: per window, mean over its multimodal ground truths of the min-over-samples ADE."""
    out = torch.zeros(pred.shape[0])
    for i in range(pred.shape[0]):
        p = pred[i].reshape(pred.shape[1], pred.shape[2], -1).unsqueeze(0)
        gt = mm_gt[i].reshape(mm_gt[i].shape[0], pred.shape[2], -1).unsqueeze(1)
        out[i] = torch.linalg.norm(p - gt, dim=-1).mean(-1).min(-1).values.mean()
    return out


def mmfde(pred: Tensor, mm_gt) -> Tensor:
    """src/metrics/multimodal.py:122-135: as mmade with the last frame's distance."""
    out = torch.zeros(pred.shape[0])
    for i in range(pred.shape[0]):
        p = pred[i].reshape(pred.shape[1], pred.shape[2], -1).unsqueeze(0)
        gt = mm_gt[i].reshape(mm_gt[i].shape[0], pred.shape[2], -1).unsqueeze(1)
        out[i] = torch.linalg.norm(p - gt, dim=-1)[..., -1].min(-1).values.mean()
    return out
