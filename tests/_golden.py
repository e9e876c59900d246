"""Helpers shared by the tests: load a golden case (tests/golden/*.npz, produced by the reference
itself through tests/golden/make_golden.py) and rebuild the matching synthetic weights."""
import os

import numpy as np
import torch

import skeletondiffusion_b200 as sdb
from skeletondiffusion_b200.testing import synth_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DATASET_CASES = ["amass_perturbed", "amass_init", "amass_perturbed_iso", "h36m_perturbed", "freeman_perturbed"]
README_CASES = ["readme_perturbed", "readme_init"]
GAIN = 2.5


def load_npz(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = {}
    for k in z.files:
        a = z[k]
        out[k] = torch.from_numpy(a) if a.dtype.kind in "fiu" and a.ndim > 0 else a.item() if a.ndim == 0 else a
    return out


def tables_of(case):
    return {k[4:]: v for k, v in case.items() if k.startswith("tab_")}


def dataset_models(case, device="cpu", precision="fp32"):
    """(spec, autoencoder, diffusion, ae_sd, diff_sd) with the case's weights loaded into OUR modules."""
    spec = sdb.get_skeleton(str(case["dataset"]))
    ae, diff = sdb.build_models(spec, "cpu", if_run_as_isotropic=bool(case["iso"]), precision=precision, seed=1234 + int(case["seed"]))
    mode = str(case["mode"])
    diff_sd = synth_state_dict(diff.state_dict(), seed=int(case["seed"]), mode=mode, gain=GAIN)
    ae_sd = synth_state_dict(ae.state_dict(), seed=int(case["seed"]) + 1, mode=mode, gain=GAIN)
    for k, v in tables_of(case).items():        # the reference's own buffers (eigenvector signs)
        diff_sd[k] = v.clone()
    diff.load_state_dict(diff_sd, strict=True)
    ae.load_state_dict(ae_sd, strict=True)
    return spec, ae.to(device).eval(), diff.to(device).eval(), ae_sd, diff_sd


def dataset_cfg(spec):
    return dict(dim=96, cond_dim=96, depth=4, attn_heads=8, attn_dim_head=32, node_types=spec.nodes_type_id,
                learn_influence=True, enc_num_layers=spec.enc_num_layers)


def readme_models(case, device="cpu", precision="fp32"):
    N = 16
    corr = case["corr"]
    tabs = tables_of(case)
    model = sdb.Denoiser(dim=96, cond_dim=0, out_dim=96, channels=N, num_nodes=N)
    diff = sdb.NonisotropicGaussianDiffusion(Sigma_N=tabs["Sigma_N"], Lambda_N=tabs["Lambda_N"], U=tabs["U"], model=model,
                                             timesteps=10, precision=precision)
    if str(case["mode"]) == "init":
        sd = {k[2:]: v.clone() for k, v in case.items() if k.startswith("w_")}
        sd.update({k: v.clone() for k, v in tabs.items()})
    else:
        sd = synth_state_dict(diff.state_dict(), seed=int(case["seed"]), mode="perturbed", gain=float(case["gain"]))
        sd.update({k: v.clone() for k, v in tabs.items()})
    diff.load_state_dict(sd, strict=True)
    return diff.to(device).eval(), sd, corr


README_CFG = dict(dim=96, cond_dim=0, depth=1, attn_heads=4, attn_dim_head=32, node_types=None, learn_influence=False)


def rel_err(a, b):
    """max |a-b| / max |b|  (the 'relative' of the <=1e-4 fp32 gate: relative to the tensor's scale)."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def elementwise_err(a, b, floor=1e-3):
    """Element-wise statistic next to rel_err: max and 99.9th percentile of |a-b| / max(|b|, floor).  The floor keeps elements
    near zero (tanh-bounded latents and poses cross zero) from dividing a 1e-7 rounding difference by 1e-9."""
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    e = (a - b).abs() / b.abs().clamp_min(floor)
    k = max(1, int(0.999 * e.numel()))
    return float(e.max()), float(e.kthvalue(k).values)
