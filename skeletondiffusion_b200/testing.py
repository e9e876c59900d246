"""Deterministic synthetic weights shared by the golden-vector generator, the tests and bench.py.

The reference's random initialisation is nearly vacuous as a parity test: typed
weights are ~0.005 in magnitude, every node type starts identical, G = I, G_add = 0 and the
clamp never fires.  `synth_state_dict(..., mode="perturbed")` therefore draws O(1/sqrt(fan_in))
weights that differ per node type and dense, non-identity influence matrices.  Values come from
numpy's frozen `RandomState` stream keyed by (seed, parameter name), so the generator script
(run where /root/reference exists) and the tests (run on the GPU box) build identical tensors
without shipping 130 MB of weights.
"""
from __future__ import annotations

import zlib
from typing import Dict

import numpy as np
import torch

__all__ = ["synth_state_dict", "synth_tensor"]

_TABLE_KEYS = ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "Lambda_N", "Sigma_N", "U",
               "U_transposed", "inv_sqrt_Lambda_bar", "Umm_sqrt_Lambda_bar", "Lambda_posterior", "posterior_mean_coef",
               "mahalanobis_S_sqrt_recip", "loss_weight")


def synth_tensor(name: str, shape, seed: int, std: float = 1.0) -> torch.Tensor:
    rs = np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return torch.from_numpy((rs.standard_normal(tuple(shape)) * std).astype(np.float32))


def synth_state_dict(reference_sd: Dict[str, torch.Tensor], seed: int = 0, mode: str = "perturbed", gain: float = 1.0,
                     g_noise: float = 0.2) -> Dict[str, torch.Tensor]:
    """Returns a full state_dict with the keys/shapes of `reference_sd`.

    mode "perturbed": every learnable tensor is regenerated (see module docstring);
    mode "init":      tensors are returned unchanged (reference-style random initialisation)."""
    out = {}
    for k, v in reference_sd.items():
        leaf = k.split(".")[-1]
        if mode == "init" or leaf in ("node_type_index", "phase") or any(k.startswith(t) or k == t for t in _TABLE_KEYS):
            out[k] = v.clone()
            continue
        shape = tuple(v.shape)
        if leaf == "G":
            out[k] = torch.eye(shape[0]) + synth_tensor(k, shape, seed, g_noise)
        elif leaf == "G_add":
            out[k] = synth_tensor(k, shape, seed, 0.02)
        elif leaf == "g":
            out[k] = 1.0 + synth_tensor(k, shape, seed, 0.1)
        elif leaf.startswith("bias"):
            out[k] = synth_tensor(k, shape, seed, 0.1)
        elif leaf.startswith("weight"):
            fan_in = shape[-1]
            out[k] = synth_tensor(k, shape, seed, gain / fan_in ** 0.5)
        else:
            raise KeyError(f"synth_state_dict: no rule for '{k}' {shape}")
    return out


def fake_predictor(seed: int, num_samples: int, pred_length: int):
    """Deterministic stand-in for get_prediction used to pin the control flow of the long-term evaluation against the
    reference's own function: pred[w, s, t] = tanh(mean_frames(obs[w]) * a[s] + b[s, t]) with seeded a, b."""
    import numpy as np
    rs = np.random.RandomState(seed)
    a = torch.from_numpy(rs.uniform(0.5, 1.5, size=(num_samples, 1, 1, 1)).astype("float32"))
    holder = {}

    def predict(obs: torch.Tensor, **_):
        W, _t, J, F = obs.shape
        if "b" not in holder:
            holder["b"] = torch.from_numpy(np.random.RandomState(seed + 1).normal(0, 0.4, size=(num_samples, pred_length, J, F)).astype("float32"))
        m = obs.float().mean(dim=1, keepdim=True).unsqueeze(1)                      # [W, 1, 1, J, F]
        return torch.tanh(m * a.to(obs.device).unsqueeze(0) + holder["b"].to(obs.device).unsqueeze(0)).contiguous()

    return predict


def grad_probe_positions(name: str, numel: int, count: int = 64):
    """Fixed pseudo-random positions at which the training goldens sample a large gradient tensor (tests/golden/make_training.py)."""
    import zlib
    import numpy as np
    rs = np.random.RandomState(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    return torch.from_numpy(rs.randint(0, numel, size=count).astype("int64"))
