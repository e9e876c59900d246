#!/usr/bin/env python
"""Golden generator for the training path: the UNMODIFIED reference's NonisotropicGaussianDiffusion.p_losses
(src/core/diffusion/base.py:262-300) with torch autograd on CPU, stress weights (dense graph influence, typed weights), injected
noise and steps.  Two cases per dataset configuration:

  k1   n_train_samples = 1: loss.mean().backward()                                   (README.md:91-93)
  kbest n_train_samples = k: TrainerDiffusion.loss with similarity_space = 'latent_space' (src/core/trainer.py:205-234, restated
        here line by line because the Trainer class needs ignite / a dataset to be constructed): per observation the loss of the
        sample with the smallest loss, times loss_weight[t], mean, backward.

Stored: inputs, loss, model_out and, for every parameter, the gradient's L2 norm and 64 entries at fixed pseudo-random positions
(the full gradients are ~60 MB); parameters with fewer than 4096 entries are stored whole.  tests/golden/training.npz.
    python tests/golden/make_training.py      (build container only)"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SKELDIFF_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "_stubs"), REF, ROOT]
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
import torch  # noqa: E402

import skeletondiffusion_b200 as sdb  # noqa: E402
from skeletondiffusion_b200.testing import synth_state_dict, grad_probe_positions  # noqa: E402
from src.core import DiffusionManager as RefDiffusionManager  # noqa: E402
from src.data.skeleton import create_skeleton  # noqa: E402

torch.set_num_threads(8)
ARCH = dict(depth=4, attn_heads=8, attn_dim_head=32, use_attention=True, self_condition=False, norm_type="none", learn_influence=True)
NUM_JOINTS = {"amass": 22, "h36m": 17, "freeman": 18}
GAIN = 1.5


def reference_diffusion(name, seed):
    spec = sdb.get_skeleton(name)
    sk = create_skeleton(dataset_name=name, motion_repr_type="SkeletonRescalePose", num_joints=NUM_JOINTS[name], if_consider_hip=False,
                         obs_length=spec.obs_length, pred_length=spec.pred_length, pose_box_size=spec.pose_box_size)
    mgr = RefDiffusionManager(diffusion_type="NonisotropicGaussianDiffusion", skeleton=sk, covariance_matrix_type="adjacency",
                              num_nodes=sk.num_nodes, node_types=sk.nodes_type_id, diffusion_conditioning=True, latent_size=96,
                              diffusion_timesteps=10, diffusion_objective="pred_x0", beta_schedule="cosine", diffusion_arch=dict(ARCH))
    diff = mgr.get_diffusion().train()
    _, ours = sdb.build_models(spec, "cpu", seed=1234 + seed)
    sd = synth_state_dict(ours.state_dict(), seed=seed, mode="perturbed", gain=GAIN)
    for k, v in diff.state_dict().items():              # keep the reference's own tables (eigenvector signs)
        if not k.startswith("model."):
            sd[k] = v.clone()
    diff.load_state_dict(sd, strict=True)
    return spec, diff, {k: v.clone() for k, v in sd.items() if not k.startswith("model.")}


def record(prefix, diff, out):
    for name, p in diff.model.named_parameters():
        g = p.grad.detach().reshape(-1)
        out[f"{prefix}.norm.{name}"] = np.float64(g.double().norm().item())
        if g.numel() <= 4096:
            out[f"{prefix}.full.{name}"] = g.numpy().copy()
        else:
            out[f"{prefix}.probe.{name}"] = g[grad_probe_positions(name, g.numel())].numpy().copy()
        p.grad = None


out = {}
for name, seed, B, k in (("amass", 31, 6, 5), ("h36m", 32, 5, 4)):
    spec, diff, tables = reference_diffusion(name, seed)
    N = spec.num_nodes
    out.update({f"{name}.tab.{k}": v.numpy() for k, v in tables.items()})       # the reference's own buffers (LAPACK eigenvector signs)
    g = torch.Generator().manual_seed(seed)
    x_start = torch.tanh(torch.randn(B, N, 96, generator=g))
    x_cond = torch.tanh(torch.randn(B, N, 96, generator=g))
    t = torch.randint(0, 10, (B,), generator=g)
    noise1 = torch.randn(B, N, 96, generator=g)
    noisek = torch.randn(B * k, N, 96, generator=g)
    pre = name
    out.update({f"{pre}.x_start": x_start.numpy(), f"{pre}.x_cond": x_cond.numpy(), f"{pre}.t": t.numpy(), f"{pre}.noise1": noise1.numpy(),
                f"{pre}.noisek": noisek.numpy(), f"{pre}.k": k, f"{pre}.seed": seed, f"{pre}.gain": GAIN})
    # ---- k = 1
    loss, w, model_out = diff.p_losses(x_start, t, noise=noise1, x_cond=x_cond, n_train_samples=1)
    loss.mean().backward()
    out[f"{pre}.k1.loss"], out[f"{pre}.k1.weight"], out[f"{pre}.k1.model_out"] = loss.detach().numpy(), w.numpy(), model_out.detach().numpy()
    record(f"{pre}.k1", diff, out)
    # ---- best of k (trainer.py:224-234 with similarity_space == 'latent_space': :214-221)
    loss, w, model_out = diff.p_losses(x_start, t, noise=noisek, x_cond=x_cond, n_train_samples=k)
    with torch.no_grad():
        closest = loss.view(B, -1).min(axis=-1).indices                                            # trainer.py:218
    sim_loss = torch.gather(loss.view(B, -1), dim=1, index=closest.unsqueeze(1)).squeeze(-1)       # :219
    sim_loss = sim_loss * w                                                                        # :232
    sim_loss.mean().backward()                                                                     # :234
    out[f"{pre}.kbest.loss"], out[f"{pre}.kbest.closest"], out[f"{pre}.kbest.total"] = loss.detach().numpy(), closest.numpy(), sim_loss.mean().item()
    record(f"{pre}.kbest", diff, out)
    print(name, "k1 loss", out[f"{pre}.k1.loss"][:3], "kbest total", out[f"{pre}.kbest.total"])

path = os.path.join(HERE, "training.npz")
np.savez_compressed(path, **out)
print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB, {len(out)} arrays)")
