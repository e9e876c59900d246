"""Training path: forward() / p_losses with autograd through the CUDA backward kernels, and the best-of-k
relaxation of TrainerDiffusion.loss.  Checked against (a) torch autograd over the float64 oracle for every differentiable op and
(b) gradients produced by the reference's own p_losses + autograd (tests/golden/training.npz, made by tests/golden/make_training.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import skeldiff_oracle as oc
from tests import _golden as G

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("dataset,kin,kout,dense", [("amass", 192, 192, True), ("amass", 192, 768, True), ("h36m", 256, 192, False), ("freeman", 96, 96, True)])
def test_graph_linear_gradients_vs_float64_autograd(cuda_device, dataset, kin, kout, dense):
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200 import training
    from skeletondiffusion_b200.testing import synth_state_dict
    spec = sdb.get_skeleton(dataset)
    N, nt = spec.num_nodes, spec.nodes_type_id
    layer = sdb.StaticGraphLinear(kin, kout, bias=True, num_nodes=N, node_types=nt, learn_influence=True)
    sd = synth_state_dict(layer.state_dict(), seed=5, mode="perturbed", gain=1.0)
    if not dense:
        sd["G"] = torch.eye(N)
    layer.load_state_dict(sd)
    g = torch.Generator().manual_seed(kin + kout)
    B = 37
    x = torch.randn(B, N, kin, generator=g)
    dout = torch.randn(B, N, kout, generator=g)
    # float64 truth through the oracle's statement of graph_structural.py:30-43
    sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    x64 = x.double().requires_grad_(True)
    out64 = oc.graph_linear(sd64, "", x64, nt, True)
    out64.backward(dout.double())
    d = cuda_device
    layer = layer.to(d)
    xd = x.to(d).requires_grad_(True)
    out = training.graph_linear(layer, xd)
    out.backward(dout.to(d))
    assert _rel(out.detach().cpu(), out64.detach()) < 3e-6
    assert _rel(xd.grad.cpu(), x64.grad) < 5e-6
    assert _rel(layer.weight.grad.cpu(), sd64["weight"].grad) < 5e-6
    assert _rel(layer.bias.grad.cpu(), sd64["bias"].grad) < 5e-6
    assert _rel(layer.G.grad.cpu(), sd64["G"].grad) < 2e-5
    # bitwise repeatable (fixed-order reductions)
    w1 = layer.weight.grad.clone(); g1 = layer.G.grad.clone()
    layer.zero_grad()
    training.graph_linear(layer, xd).backward(dout.to(d))
    assert torch.equal(w1, layer.weight.grad) and torch.equal(g1, layer.G.grad)


def test_block_rmsnorm_attention_loss_gradients(cuda_device):
    from skeletondiffusion_b200 import training
    d = cuda_device
    g = torch.Generator().manual_seed(3)
    B, N, C, T = 19, 21, 192, 10
    # scale / shift / tanh with a per-step table
    y = torch.randn(B, N, C, generator=g); tab = torch.randn(T, 2 * C, generator=g) * 0.3; t = torch.randint(0, T, (B,), generator=g)
    dh = torch.randn(B, N, C, generator=g)
    y64, tab64 = y.double().requires_grad_(True), tab.double().requires_grad_(True)
    sc, sh = tab64[t][:, None, :C], tab64[t][:, None, C:]
    torch.tanh(y64 * (sc + 1) + sh).backward(dh.double())
    yd, tabd = y.to(d).requires_grad_(True), tab.to(d).requires_grad_(True)
    h = training._SsTanhFn.apply(yd, tabd, t.to(d, torch.int32))
    h.backward(dh.to(d))
    assert _rel(yd.grad.cpu(), y64.grad) < 5e-6 and _rel(tabd.grad.cpu(), tab64.grad) < 5e-6
    # RMSNorm
    x = torch.randn(B, N, C, generator=g); gg = torch.rand(1, 1, C, generator=g) + 0.5
    x64, g64 = x.double().requires_grad_(True), gg.double().requires_grad_(True)
    (torch.nn.functional.normalize(x64, dim=-1) * g64 * C ** 0.5).backward(dh.double())
    xd, gd = x.to(d).requires_grad_(True), gg.to(d).requires_grad_(True)
    training._RMSNormFn.apply(xd, gd).backward(dh.to(d))
    assert _rel(xd.grad.cpu(), x64.grad) < 5e-6 and _rel(gd.grad.cpu(), g64.grad) < 5e-6
    # node attention, 8 heads x 32
    qkv = torch.randn(B, N, 768, generator=g); do = torch.randn(B, N, 256, generator=g)
    q64 = qkv.double().requires_grad_(True)
    q, k, v = [u.reshape(B, N, 8, 32).permute(0, 2, 1, 3) for u in q64.chunk(3, -1)]
    att = torch.einsum("bhnc,bhjc->bhnj", q * 32 ** -0.5, k).softmax(-1)
    o64 = torch.einsum("bhnj,bhjd->bhnd", att, v).permute(0, 2, 1, 3).reshape(B, N, 256)
    o64.backward(do.double())
    qd = qkv.to(d).requires_grad_(True)
    o = training._NodeAttentionFn.apply(qd, 8, 32)
    o.backward(do.to(d))
    assert _rel(o.detach().cpu(), o64.detach()) < 3e-6 and _rel(qd.grad.cpu(), q64.grad) < 5e-6
    # Mahalanobis L1 loss
    S = torch.randn(T, N, N, generator=g); out = torch.randn(B, N, 96, generator=g); x0 = torch.randn(B, N, 96, generator=g); gl = torch.rand(B, generator=g)
    o64 = out.double().requires_grad_(True)
    ((S.double()[t] @ (o64 - x0.double())).abs().mean((1, 2))).backward(gl.double())
    od = out.to(d).requires_grad_(True)
    training._MahalanobisL1Fn.apply(od, x0.to(d), t.to(d, torch.int32), S.to(d)).backward(gl.to(d))
    assert _rel(od.grad.cpu(), o64.grad) < 5e-6


def _training_case(name, device):
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200.testing import synth_state_dict
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "training.npz"))
    spec = sdb.get_skeleton(name)
    seed = int(z[f"{name}.seed"])
    _, diff = sdb.build_models(spec, "cpu", seed=1234 + seed)
    sd = synth_state_dict(diff.state_dict(), seed=seed, mode="perturbed", gain=float(z[f"{name}.gain"]))
    for key in z.files:                                   # the reference's own diffusion buffers (LAPACK eigenvector signs)
        if key.startswith(f"{name}.tab."):
            sd[key[len(name) + 5:]] = torch.from_numpy(z[key])
    diff.load_state_dict(sd, strict=True)
    diff = diff.to(device).train()
    get = lambda key: torch.from_numpy(z[f"{name}.{key}"]).to(device)
    return z, diff, get


def _check_grads(z, prefix, diff, tol):
    from skeletondiffusion_b200.testing import grad_probe_positions
    worst = 0.0
    for pname, p in diff.model.named_parameters():
        assert p.grad is not None, pname
        g = p.grad.detach().reshape(-1).cpu()
        ref_norm = float(z[f"{prefix}.norm.{pname}"])
        if f"{prefix}.full.{pname}" in z.files:
            ref, got = torch.from_numpy(z[f"{prefix}.full.{pname}"]), g
        else:
            ref, got = torch.from_numpy(z[f"{prefix}.probe.{pname}"]), g[grad_probe_positions(pname, g.numel())]
        scale = max(ref_norm / max(g.numel(), 1) ** 0.5, 1e-12)          # rms magnitude of the reference gradient
        err = float((got.double() - ref.double()).abs().max()) / max(float(ref.abs().max()), scale)
        worst = max(worst, err)
        assert err < tol, (pname, err)
        assert abs(float(g.double().norm()) - ref_norm) <= tol * max(ref_norm, 1e-12), (pname, float(g.double().norm()), ref_norm)
    return worst


@pytest.mark.parametrize("name", ["amass", "h36m"])
def test_p_losses_gradients_match_the_reference(cuda_device, name):
    """n_train_samples = 1: loss values, model output and every parameter gradient against the reference's own autograd."""
    z, diff, get = _training_case(name, cuda_device)
    loss, w, out = diff.p_losses(get("x_start"), get("t"), noise=get("noise1"), x_cond=get("x_cond"), n_train_samples=1)
    assert _rel(loss.detach().cpu(), torch.from_numpy(z[f"{name}.k1.loss"])) < 1e-4
    assert _rel(out.detach().cpu(), torch.from_numpy(z[f"{name}.k1.model_out"])) < 1e-4
    assert torch.allclose(w.cpu(), torch.from_numpy(z[f"{name}.k1.weight"]))
    loss.mean().backward()
    worst = _check_grads(z, f"{name}.k1", diff, 1e-4)
    print(f"{name}: worst parameter-gradient error vs the reference (k = 1): {worst:.1e}")


@pytest.mark.parametrize("name", ["amass", "h36m"])
def test_best_of_k_training_step_matches_the_reference(cuda_device, name):
    """TrainerDiffusion.loss with k samples per observation (trainer.py:205-234): the sparse-row backward must give the reference's
    gradients, and must equal the dense autograd through all B * k rows."""
    from skeletondiffusion_b200 import training
    z, diff, get = _training_case(name, cuda_device)
    k, B = int(z[f"{name}.k"]), get("x_start").shape[0]
    loss, w, _ = diff.p_losses(get("x_start"), get("t"), noise=get("noisek"), x_cond=get("x_cond"), n_train_samples=k)
    assert _rel(loss.detach().cpu(), torch.from_numpy(z[f"{name}.kbest.loss"])) < 1e-4
    sim, idx = training.ksimilarity_loss(loss, B)
    assert torch.equal(idx.cpu(), torch.from_numpy(z[f"{name}.kbest.closest"]))
    total = (sim * w).mean()
    assert abs(float(total.detach()) - float(z[f"{name}.kbest.total"])) < 1e-5
    total.backward()
    worst = _check_grads(z, f"{name}.kbest", diff, 1e-4)
    sparse = {n: p.grad.clone() for n, p in diff.model.named_parameters()}
    # dense autograd over all rows
    diff.zero_grad()
    xs, t = get("x_start").repeat_interleave(k, 0), get("t").repeat_interleave(k, 0)
    xn = diff.q_sample(xs, t, get("noisek"))
    dl, _ = training.diffusion_loss_train(diff, xn, xs, t, get("x_cond").repeat_interleave(k, 0))
    (torch.gather(dl.view(B, -1), 1, idx.unsqueeze(1)).squeeze(-1) * w).mean().backward()
    for n, p in diff.model.named_parameters():
        assert _rel(sparse[n].cpu(), p.grad.cpu()) < 2e-5 or float(p.grad.abs().max()) < 1e-12, n
    print(f"{name}: worst parameter-gradient error vs the reference (best of {k}): {worst:.1e}")


def test_forward_trains(cuda_device):
    """README usage (README.md:88-93): loss, _, _ = diffusion(training_samples); loss.mean().backward(); a few Adam steps on a
    fixed batch lower the loss."""
    import skeletondiffusion_b200 as sdb
    torch.manual_seed(0)
    N = 16
    rand = (torch.rand(N, N) >= 0.5).float()
    corr = (rand + rand.T) // 2
    S_, L_, U_ = sdb.get_cov_from_corr(corr, if_sigma_n_scale=True, sigma_n_scale="spectral")
    model = sdb.Denoiser(dim=96, cond_dim=0, out_dim=96, channels=N, num_nodes=N)
    diff = sdb.NonisotropicGaussianDiffusion(Sigma_N=S_, Lambda_N=L_, U=U_, model=model, timesteps=10).to(cuda_device).train()
    x = torch.rand(8, N, 96, device=cuda_device)
    loss, _, _ = diff(x)
    loss.mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters() if p.requires_grad)
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    t = torch.randint(0, 10, (8,), device=cuda_device)
    noise = torch.randn(8, N, 96, device=cuda_device)
    hist = []
    for _ in range(25):
        opt.zero_grad()
        l, _, _ = diff.p_losses(x, t, noise=noise)
        l.mean().backward()
        opt.step()
        hist.append(float(l.mean()))
    assert hist[-1] < 0.7 * hist[0], hist
