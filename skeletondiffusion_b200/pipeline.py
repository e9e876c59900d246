"""Caller-side glue of the sampling path with the reference's signatures.

Reference: get_diffusion_latent_codes / decode_latent_pred / get_prediction
(src/eval_prepare_model.py:89-121) and DiffusionManager (src/core/diffusion_manager.py:8-45).
The reference materialises `obs` and `z_past` num_samples times with repeat_interleave; here
the kernels read the per-window rows in place (sd_view.rep), so nothing is replicated.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import torch

from .autoencoder import AutoEncoder
from .diffusion import NonisotropicGaussianDiffusion, get_cov_from_corr
from .network import Denoiser

from .plan import params_key

__all__ = ["DiffusionManager", "get_diffusion_latent_codes", "decode_latent_pred", "get_prediction", "GraphedPrediction",
           "load_model_checkpoint", "prepare_model",
           "best_sample", "long_term_prediction_best_every50", "shard_windows", "build_models"]


class DiffusionManager:
    """Factory with the reference's keyword surface (diffusion_manager.py:8-45)."""

    def __init__(self, diffusion_type: str = "IsotropicGaussianDiffusion", skeleton=None, covariance_matrix_type: str = "adjacency",
                 reachability_matrix_degree_factor=0.5, reachability_matrix_stop_at=0, if_sigma_n_scale=True,
                 sigma_n_scale="spectral", if_run_as_isotropic=False, **kwargs):
        model = self.get_network(**kwargs)
        self.diffusion_type = diffusion_type
        if diffusion_type != "NonisotropicGaussianDiffusion":
            raise NotImplementedError("IsotropicGaussianDiffusion is the reference's ablation class (SURVEY §2 row 4: out of scope); "
                                      "use NonisotropicGaussianDiffusion with if_run_as_isotropic=True")
        if covariance_matrix_type == "adjacency":
            corr = skeleton.adj_matrix
        elif covariance_matrix_type == "reachability":
            corr = skeleton.reachability_matrix(factor=reachability_matrix_degree_factor, stop_at=reachability_matrix_stop_at)
        else:
            raise AssertionError("Not implemented")
        sigma, lam, u = get_cov_from_corr(correlation_matrix=corr, if_sigma_n_scale=if_sigma_n_scale, sigma_n_scale=sigma_n_scale,
                                          if_run_as_isotropic=if_run_as_isotropic, **kwargs)
        self.diffusion = NonisotropicGaussianDiffusion(Sigma_N=sigma, Lambda_N=lam, U=u, model=model, **kwargs)

    def get_diffusion(self):
        return self.diffusion

    def get_network(self, num_nodes, diffusion_conditioning=False, latent_size=96, node_types: torch.Tensor = None,
                    diffusion_arch: Dict[str, Any] = None, **kwargs):
        cond_dim = latent_size if diffusion_conditioning else 0
        arch = dict(diffusion_arch or {})
        arch.pop("arch", None)
        return Denoiser(dim=latent_size, cond_dim=cond_dim, out_dim=latent_size, channels=num_nodes, num_nodes=num_nodes,
                        node_types=node_types, **arch)


def load_model_checkpoint(load_path: str, map_location="cpu"):
    """src/utils/load.py:11-17: the reference's checkpoints are `torch.save`d dicts whose 'model' entry is the state_dict
    (the other entries - optimizer, epoch, EMA - belong to the trainer and are ignored here)."""
    return torch.load(load_path, map_location=map_location, weights_only=False)


def prepare_model(skeleton, autoencoder_checkpoint: str, diffusion_checkpoint: str, device="cuda", precision: str = "fp16x2", **manager_kwargs):
    """prepare_autoencoder + prepare_model (src/eval_prepare_model.py:26-85) for the published layout: build the dataset-config
    modules, `load_state_dict(checkpoint['model'])` with strict keys (the module and buffer names are the reference's, Appendix B
    of the survey), move to `device`, eval().  Returns ((autoencoder, diffusion), device)."""
    ae, diffusion = build_models(skeleton, "cpu", precision=precision, seed=None, **manager_kwargs)
    ae.load_state_dict(load_model_checkpoint(autoencoder_checkpoint)["model"], strict=True)
    diffusion.load_state_dict(load_model_checkpoint(diffusion_checkpoint)["model"], strict=True)
    device = torch.device(device)
    return (ae.to(device).eval(), diffusion.to(device).eval()), device


def get_diffusion_latent_codes(obs, model, num_samples=50, **kwargs):
    autoencoder, diffusion = model
    bs = obs.shape[0]
    sampler_kwargs = kwargs.get("sampler_kwargs", {})
    # the autoencoder's recurrent products follow the diffusion's precision mode ('bf16x3' is fp32-grade on the tensor cores)
    prec = getattr(diffusion, "precision", None)
    past_embedding = autoencoder.get_past_embedding(obs, precision=None if prec in (None, "bf16") else prec)
    if kwargs.get("diffusion_conditioning", True):
        # x_cond has bs rows, the batch bs*num_samples: rows are repeat_interleave'd in place (base.py:246-248)
        latent_pred, _ = diffusion.sample(batch_size=bs * num_samples, x_cond=past_embedding, **sampler_kwargs)
    else:
        latent_pred, _ = diffusion.sample(batch_size=bs * num_samples, **sampler_kwargs)
    return latent_pred, past_embedding


def decode_latent_pred(obs, latent_pred, z_past, model, num_samples=50, pred_length=100, **kwargs):
    autoencoder, diffusion = model
    bs, _, j, f = obs.shape
    prec = getattr(diffusion, "precision", None)
    pred = autoencoder.decode(obs, latent_pred, z_past, ph=pred_length,        # obs rows shared by the samples of a window
                              precision=None if prec is None else ("bf16x3" if prec == "bf16" else prec))
    return pred.view(bs, num_samples, pred_length, j, f)


def get_prediction(obs, model, num_samples=50, pred_length=100, **kwargs):
    """obs [W, T_obs, N, 3] -> predictions [W, num_samples, pred_length, N, 3] (eval_prepare_model.py:118-121)."""
    lat_pred, z_past = get_diffusion_latent_codes(obs, model, num_samples=num_samples, **kwargs)
    return decode_latent_pred(obs, lat_pred, z_past, model, num_samples=num_samples, pred_length=pred_length, **kwargs)


def best_sample(pred: torch.Tensor, target: torch.Tensor, keep_frames: int = 0, scale: float = 1.0):
    """get_best_sample_idx (src/metrics/utils.py:22-30) on the device: per window the sample closest to `target` in mean per-joint
    distance.  pred [W, S, T, J, 3], target [W, T, J, 3] -> (best [W, T, J, 3], tail [W, keep_frames, J, 3], index [W]),
    best / tail multiplied by `scale`."""
    from . import _native as nv
    nv.require_cuda(pred, "pred")
    W, S, T, J, F = pred.shape
    assert F == 3 and tuple(target.shape) == (W, T, J, 3), f"target {tuple(target.shape)} does not match pred {tuple(pred.shape)}"
    p, tg = pred.float().contiguous(), target.to(pred.device, torch.float32).contiguous()
    best = torch.empty(W, T, J, 3, device=p.device)
    tail = torch.empty(W, keep_frames, J, 3, device=p.device)
    index = torch.empty(W, device=p.device, dtype=torch.int32)
    nv.check(nv.load().sd_best_sample(p.data_ptr(), tg.data_ptr(), W, S, T, J, keep_frames, float(scale), best.data_ptr(),
                                      tail.data_ptr() if keep_frames else None, index.data_ptr(), nv.stream_ptr(p.device)), "sd_best_sample")
    return best, tail, index


@torch.no_grad()
def long_term_prediction_best_every50(data: torch.Tensor, target: torch.Tensor, model, skeleton, num_samples: int = 50,
                                      pred_length: int = 100, long_term_factor: float = 2.0, sampler_kwargs_per_segment=None,
                                      predict_fn=None):
    """Long-term autoregressive evaluation (src/eval_utils.py:44-67): predict `num_samples` continuations, keep per window the
    one closest to the ground-truth segment, feed its last obs-length frames back in as the next observation, and so on for
    ceil(long_term_factor) segments (the last one cut to the fractional part).  Everything stays on the device: the
    reference's per-segment `.cpu()` of the arg-min indices (src/metrics/utils.py:24) is the sd_best_sample kernel.
    Like the reference (eval_utils.py:58-65 after process_evaluation_pair), the frames fed back are the METRIC-space ones
    (multiplied by pose_box_size); this is reproduced, not corrected.
    Returns (target [W, L, J, 3], pred [W, num_samples, L, J, 3] (the chosen continuation repeated), obs in metric space).
    NB the reference function needs `math` in its module namespace (src/eval_utils.py never imports it)."""
    import math
    n_seg = math.ceil(long_term_factor)
    n_past = data.shape[-3]
    scale = float(skeleton.pose_box_size)
    new_data, finals, targets = data, [], []
    for idx in range(n_seg):
        kw = {} if sampler_kwargs_per_segment is None else dict(sampler_kwargs=sampler_kwargs_per_segment[idx])
        if predict_fn is not None:                 # the reference takes the predictor as a callable as well (eval.py:73-74)
            pred = predict_fn(new_data)
        else:
            pred = get_prediction(new_data, model, num_samples=num_samples, pred_length=pred_length, diffusion_conditioning=True, **kw)
        if idx == n_seg - 1 and int(long_term_factor) != long_term_factor:
            pred = pred[..., :int(long_term_factor * pred_length) % pred_length, :, :].contiguous()
        seg = target[..., idx * pred_length:(idx + 1) * pred_length, :, :][..., :pred.shape[2], :, :]
        best, tail, _ = best_sample(pred, seg.to(pred.device), keep_frames=min(n_past, pred.shape[2]), scale=scale)
        finals.append(best)
        targets.append(seg.to(pred.device) * scale)
        new_data = tail
    final = torch.cat(finals, dim=-3)
    return torch.cat(targets, dim=-3), final.unsqueeze(1).repeat(1, num_samples, 1, 1, 1), data * scale


class GraphedPrediction:
    """`get_prediction` for a fixed (windows, num_samples, pred_length) captured ONCE as one CUDA graph: encode, the T-step
    sampling loop (T x (Denoiser kernels + reverse step)) and the decode of every frame -- about 1 100 kernel launches for the
    AMASS configuration -- are replayed with a single cudaGraphLaunch (north_star item 3; the reference's caller is
    src/eval_prepare_model.py:118-121).  Fresh noise is drawn by two Philox launches in front of every replay (or injected);
    observations are copied into the graph's static input.  The capture holds references to the packed plans it baked in and
    is rebuilt when a parameter or buffer of the models changes (load_state_dict, optimizer step, .to())."""

    def __init__(self, model, windows: int, num_samples: int = 50, pred_length: int = 100, device=None):
        self.model, self.W, self.S, self.ph = model, int(windows), int(num_samples), int(pred_length)
        ae, diff = model
        self.device = torch.device(device) if device is not None else diff.betas.device
        self._graph = None
        self._key = None

    def _models_key(self):
        ae, diff = self.model
        return (params_key(list(ae.parameters()) + list(ae.buffers()) + list(diff.parameters()) + list(diff.buffers())), diff.precision)

    def _capture(self, obs_shape):
        ae, diff = self.model
        dev = self.device
        N, D, T = diff.channels, diff.seq_length, diff.num_timesteps
        B = self.W * self.S
        st = dict(obs=torch.zeros(obs_shape, device=dev, dtype=torch.float32),
                  start=torch.zeros(B, N, D, device=dev, dtype=torch.float32),
                  noise=torch.zeros(B, max(T - 1, 1), N, D, device=dev, dtype=torch.float32)[:, :T - 1])

        def run():
            return get_prediction(st["obs"], self.model, num_samples=self.S, pred_length=self.ph, diffusion_conditioning=True,
                                  sampler_kwargs=dict(start_noise=st["start"], sampling_noise=st["noise"] if T > 1 else None))

        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                  # warm-up outside the capture: plans, workspaces, function attributes
            run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            st["pred"] = run()
        self._graph, self._static, self._key = graph, st, self._models_key()
        self._plans = (ae, diff, diff.model.plan(), diff._diffusion_plan())      # keep what the graph baked in alive

    @torch.no_grad()
    def __call__(self, obs: torch.Tensor, start_noise: Optional[torch.Tensor] = None, sampling_noise: Optional[torch.Tensor] = None,
                 clone: bool = False, window_offset: int = 0) -> torch.Tensor:
        """obs [W, T_obs, N, 3] -> predictions [W, S, pred_length, N, 3] (the graph's static output buffer unless clone=True).
        window_offset: global index of obs[0] (a rank's shard / a chunk of a larger job): the noise of a window is drawn from the
        Philox stream at its GLOBAL row index, so results do not depend on the number of ranks or the chunking."""
        ae, diff = self.model
        if obs.shape[0] != self.W:
            raise ValueError(f"graph was built for {self.W} windows, got {obs.shape[0]}")
        if self._graph is None or self._key != self._models_key() or tuple(self._static["obs"].shape) != tuple(obs.shape):
            self._graph = None
            self._capture(tuple(obs.shape))
        st = self._static
        st["obs"].copy_(obs, non_blocking=True)
        if start_noise is not None:
            st["start"].copy_(start_noise)
        else:
            diff.fill_noise_(st["start"], offset=int(window_offset) * self.S * st["start"][0].numel())
        if st["noise"].numel():
            if sampling_noise is not None:
                st["noise"].copy_(sampling_noise)
            else:
                diff.fill_noise_(st["noise"], offset=int(window_offset) * self.S * st["noise"][0].numel())
        self._graph.replay()
        return st["pred"].clone() if clone else st["pred"]


def shard_windows(num_windows: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous window range [lo, hi) of `rank`; all samples of a window stay on one GPU so APD/ADE/FDE
    (which reduce over the samples of a window) need no exchange."""
    base, rem = divmod(num_windows, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def build_models(skeleton, device, diffusion_timesteps: int = 10, depth: int = 4, attn_heads: int = 8, attn_dim_head: int = 32,
                 latent_size: int = 96, if_run_as_isotropic: bool = False, precision: str = "fp32", seed: Optional[int] = 0):
    """Dataset-config models (configs/config_train_diffusion/model/skeleton_diffusion.yaml:49-57,
    config_train_autoencoder/model/autoencoder.yaml) with reference-style random initialisation."""
    if seed is not None:
        torch.manual_seed(seed)
    nt = skeleton.nodes_type_id
    ae = AutoEncoder(num_nodes=skeleton.num_nodes, encoder_hidden_size=96, decoder_hidden_size=96, latent_size=latent_size,
                     node_types=nt, input_size=3, z_activation="tanh", enc_num_layers=skeleton.enc_num_layers,
                     recurrent_arch_enc="StaticGraphGRU", recurrent_arch_decoder="StaticGraphGRU", output_size=3,
                     if_consider_hip=False)
    mgr = DiffusionManager(diffusion_type="NonisotropicGaussianDiffusion", skeleton=skeleton, covariance_matrix_type="adjacency",
                           num_nodes=skeleton.num_nodes, node_types=nt, diffusion_conditioning=True, latent_size=latent_size,
                           diffusion_timesteps=diffusion_timesteps, diffusion_objective="pred_x0", beta_schedule="cosine",
                           if_run_as_isotropic=if_run_as_isotropic, precision=precision,
                           diffusion_arch=dict(depth=depth, attn_heads=attn_heads, attn_dim_head=attn_dim_head, use_attention=True,
                                               self_condition=False, norm_type="none", learn_influence=True))
    diffusion = mgr.get_diffusion()
    return ae.to(device).eval(), diffusion.to(device).eval()
