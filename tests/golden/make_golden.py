#!/usr/bin/env python
"""Golden-vector generator: runs the UNMODIFIED reference (/root/reference) on CPU and stores
small input/output fixtures under tests/golden/*.npz.

Run only in the build container (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py
Needs three import stubs (tests/golden/_stubs): denoising_diffusion_pytorch (SinusoidalPosEmb),
hydra, omegaconf — none of them is installed here.

Weights are NOT stored: both sides regenerate them with skeletondiffusion_b200.testing.synth_state_dict
(frozen numpy RandomState stream).  Loading our state_dict into the reference modules with
strict=True doubles as the state_dict-compatibility check of the drop-in classes.
"""
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
from skeletondiffusion_b200.testing import synth_state_dict, synth_tensor  # noqa: E402

from src.core import AutoEncoder as RefAutoEncoder, DiffusionManager as RefDiffusionManager  # noqa: E402
from src.core.diffusion import NonisotropicGaussianDiffusion as RefDiffusion, get_cov_from_corr as ref_cov  # noqa: E402
from src.core.network import Denoiser as RefDenoiser  # noqa: E402
from src.data.skeleton import create_skeleton  # noqa: E402
from src.metrics.multimodal import ade as ref_ade, apd as ref_apd, fde as ref_fde  # noqa: E402
from src.eval_prepare_model import get_prediction as ref_get_prediction  # noqa: E402

torch.set_grad_enabled(False)
torch.set_num_threads(8)
ARCH = dict(depth=4, attn_heads=8, attn_dim_head=32, use_attention=True, self_condition=False, norm_type="none", learn_influence=True)
NUM_JOINTS = {"amass": 22, "h36m": 17, "freeman": 18}


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB)")


def ref_skeleton(name):
    spec = sdb.get_skeleton(name)
    sk = create_skeleton(dataset_name=name, motion_repr_type="SkeletonRescalePose", num_joints=NUM_JOINTS[name], if_consider_hip=False,
                         obs_length=spec.obs_length, pred_length=spec.pred_length, pose_box_size=spec.pose_box_size)
    # the tabulated skeleton data must agree with the reference classes
    assert sk.num_nodes == spec.num_nodes
    assert sk.nodes_type_id.tolist() == list(spec.node_types)
    assert torch.equal(sk.adj_matrix, spec.adj_matrix)
    return sk, spec


def dataset_case(name, mode, seed, windows, samples, iso=False, ph=None, with_example=False):
    sk, spec = ref_skeleton(name)
    ph = ph or spec.pred_length
    N = spec.num_nodes
    # ---- reference models
    ref_mgr = RefDiffusionManager(diffusion_type="NonisotropicGaussianDiffusion", skeleton=sk, covariance_matrix_type="adjacency",
                                  num_nodes=N, node_types=sk.nodes_type_id, diffusion_conditioning=True, latent_size=96,
                                  diffusion_timesteps=10, diffusion_objective="pred_x0", beta_schedule="cosine",
                                  if_run_as_isotropic=iso, diffusion_arch=dict(ARCH))
    ref_diff = ref_mgr.get_diffusion().eval()
    ref_ae = RefAutoEncoder(num_nodes=N, encoder_hidden_size=96, decoder_hidden_size=96, latent_size=96, node_types=sk.nodes_type_id,
                            input_size=3, z_activation="tanh", enc_num_layers=spec.enc_num_layers, recurrent_arch_enc="StaticGraphGRU",
                            recurrent_arch_decoder="StaticGraphGRU", output_size=3, if_consider_hip=False).eval()
    # ---- our modules define the state_dict; the reference must accept it strictly
    torch.manual_seed(1234 + seed)
    ae, diff = sdb.build_models(spec, "cpu", if_run_as_isotropic=iso, seed=None)
    diff_sd = synth_state_dict(diff.state_dict(), seed=seed, mode=mode, gain=2.5)
    ae_sd = synth_state_dict(ae.state_dict(), seed=seed + 1, mode=mode, gain=2.5)
    # the diffusion tables of our class must match the reference's own
    ref_tables = {k: v for k, v in ref_diff.state_dict().items() if not k.startswith("model.")}
    for k, v in ref_tables.items():
        ours = diff_sd[k]
        if k in ("U", "U_transposed", "Sigma_N", "Lambda_N") or True:
            err = (ours - v).abs().max().item()
            assert err <= 1e-6, f"table {k} differs from the reference by {err}"
        diff_sd[k] = v.clone()          # use the reference's own buffers (eigenvector signs etc.)
    ref_diff.load_state_dict(diff_sd, strict=True)
    ref_ae.load_state_dict(ae_sd, strict=True)
    # ---- inputs
    B = windows * samples
    obs = synth_tensor("obs", (windows, spec.obs_length, N, 3), seed, 0.3).clamp(-1, 1)
    if with_example:
        ex = np.load(os.path.join(REF, "figures", "example_obs_amass.npy"))[:, :, :22].astype(np.float32)
        ex = torch.from_numpy(ex)
        ex = (ex - ex[:, :, :1])[:, :, 1:] / spec.pose_box_size
        obs[0] = ex[0]
    start_noise = synth_tensor("start_noise", (B, N, 96), seed)
    sampling_noise = synth_tensor("sampling_noise", (B, 9, N, 96), seed)
    # ---- reference outputs
    z_past = ref_ae.get_past_embedding(obs)
    zc = z_past.repeat_interleave(samples, 0)
    x_probe = synth_tensor("x_probe", (B, N, 96), seed)
    t_probe = torch.tensor([(3 * i + 1) % 10 for i in range(B)], dtype=torch.long)
    den_out = ref_diff.model(x_probe, t_probe, None, zc)
    lat, (n0, noise_t, mean_t) = ref_diff.sample(batch_size=B, x_cond=zc, start_noise=start_noise.clone(),
                                                 sampling_noise=sampling_noise, return_sampling_noise=True)
    assert torch.equal(noise_t, sampling_noise)
    pred = ref_ae.decode(obs.repeat_interleave(samples, 0), lat, zc, ph=ph).view(windows, samples, ph, N, 3)
    # end-to-end through the reference's own caller with the same injected noise
    pred2 = ref_get_prediction(obs, (ref_ae, ref_diff), num_samples=samples, pred_length=ph, diffusion_conditioning=True,
                               sampler_kwargs=dict(start_noise=start_noise.clone(), sampling_noise=sampling_noise))
    assert torch.equal(pred, pred2)
    # training-loss entry point with fixed t / noise
    t_loss = torch.tensor([(7 * i + 2) % 10 for i in range(B)], dtype=torch.long)
    x_start = synth_tensor("x_start", (B, N, 96), seed, 0.5).clamp(-1, 1)
    noise_loss = synth_tensor("noise_loss", (B, N, 96), seed)
    loss, lw, mout = ref_diff.p_losses(x_start, t_loss, noise=noise_loss, x_cond=zc)
    xq = ref_diff.q_sample(x_start, t_loss, noise_loss)
    # metrics against a synthetic target
    target = synth_tensor("target", (windows, ph, N, 3), seed, 0.3).clamp(-1, 1)
    pm, tm = sk.transform_to_metric_space(pred), sk.transform_to_metric_space(target)
    frac_clamped = float((lat.abs() >= 1.0).float().mean())
    print(f"  {name}/{mode}: |x0| max {den_out.abs().max():.3f}, latents clamped {100 * frac_clamped:.2f}%, pred std {pred.std():.3f}")
    tables = {("tab_" + k): v for k, v in ref_tables.items()}
    save(f"{name}_{mode}" + ("_iso" if iso else ""), dataset=name, mode=mode, seed=seed, windows=windows, samples=samples, ph=ph, iso=int(iso),
         obs=obs, start_noise=start_noise, sampling_noise=sampling_noise, z_past=z_past, x_probe=x_probe, t_probe=t_probe,
         den_out=den_out, latents=lat, mean_t=mean_t, pred=pred, t_loss=t_loss, x_start=x_start, noise_loss=noise_loss,
         loss=loss, loss_weight=lw, loss_model_out=mout, q_sample=xq, target=target,
         ade=ref_ade(tm, pm), fde=ref_fde(tm, pm), apd=ref_apd(pm), **tables)


def readme_case(mode, seed):
    """README plug-and-play (README.md:72-98): Denoiser(dim=96, num_nodes=16), random symmetric correlation, T=10."""
    torch.manual_seed(0)
    N = 16
    ref_model = RefDenoiser(dim=96, cond_dim=0, out_dim=96, channels=N, num_nodes=N)
    rand = (torch.rand(N, N) >= 0.5).float()
    corr = (rand + rand.T) // 2
    sigma, lam, u = ref_cov(correlation_matrix=corr, if_sigma_n_scale=True, sigma_n_scale="spectral", if_run_as_isotropic=False)
    ref_diff = RefDiffusion(Sigma_N=sigma, Lambda_N=lam, U=u, model=ref_model, timesteps=10).eval()
    ours_model = sdb.Denoiser(dim=96, cond_dim=0, out_dim=96, channels=N, num_nodes=N)
    s2, l2, u2 = sdb.get_cov_from_corr(correlation_matrix=corr, if_sigma_n_scale=True, sigma_n_scale="spectral")
    assert torch.allclose(s2, sigma, atol=1e-6) and torch.allclose(l2, lam, atol=1e-6)
    ours = sdb.NonisotropicGaussianDiffusion(Sigma_N=sigma, Lambda_N=lam, U=u, model=ours_model, timesteps=10)
    # learn_influence=False: G is used un-normalised, so the per-layer gain must stay below 1 or the tanh net turns
    # chaotic (fp32 re-association noise then flips saturated signs and no implementation can match another)
    gain = 0.6
    sd = synth_state_dict(ours.state_dict(), seed=seed, mode=mode, gain=gain)
    if mode == "init":
        sd = {k: v.clone() for k, v in ref_diff.state_dict().items()}
    else:
        for k, v in ref_diff.state_dict().items():
            if not k.startswith("model."):
                assert (sd[k] - v).abs().max() <= 1e-6, k
                sd[k] = v.clone()
    assert set(sd) == set(ours.state_dict()), "state_dict keys differ"
    ref_diff.load_state_dict(sd, strict=True)
    B = 4
    start_noise = synth_tensor("start_noise", (B, N, 96), seed)
    sampling_noise = synth_tensor("sampling_noise", (B, 9, N, 96), seed)
    lat, (_, _, mean_t) = ref_diff.sample(batch_size=B, start_noise=start_noise.clone(), sampling_noise=sampling_noise, return_sampling_noise=True)
    x_probe = synth_tensor("x_probe", (B, N, 96), seed)
    t_probe = torch.tensor([9, 0, 4, 7])
    den_out = ref_diff.model(x_probe, t_probe)
    arrays = dict(corr=corr, start_noise=start_noise, sampling_noise=sampling_noise, latents=lat, mean_t=mean_t, x_probe=x_probe,
                  t_probe=t_probe, den_out=den_out, mode=mode, seed=seed, gain=gain)
    arrays.update({("tab_" + k): v for k, v in ref_diff.state_dict().items() if not k.startswith("model.")})
    if mode == "init":     # random-init weights cannot be regenerated elsewhere: store them (0.56 M parameters)
        arrays.update({("w_" + k): v for k, v in ref_diff.state_dict().items() if k.startswith("model.")})
    save(f"readme_{mode}", **arrays)


if __name__ == "__main__":
    readme_case("perturbed", seed=11)
    readme_case("init", seed=12)
    dataset_case("amass", "perturbed", seed=21, windows=2, samples=3, with_example=True)
    dataset_case("amass", "init", seed=22, windows=1, samples=2, ph=24)
    dataset_case("amass", "perturbed", seed=23, windows=1, samples=2, iso=True, ph=8)
    dataset_case("h36m", "perturbed", seed=24, windows=2, samples=2, ph=20)
    dataset_case("freeman", "perturbed", seed=25, windows=2, samples=2, ph=12)
