"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
(i) golden vectors produced by the reference's own classes and (ii) the CPU oracle on seeded inputs.
fp32 gate: max|a-b| / max|b| <= 1e-4 (BASELINE.json north_star)."""
import ctypes as C

import pytest
import torch

from oracle import skeldiff_oracle as oc
from tests import _golden as G

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4


def _native():
    from skeletondiffusion_b200 import _native as nv
    return nv


# ------------------------------------------------------------------------------------------------
# single operators against the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("typed,learn,bias,kin,kout,batch", [
    (True, True, True, 192, 192, 5), (True, True, False, 192, 768, 130), (False, False, True, 96, 96, 4),
    (True, True, True, 99, 96, 33), (True, True, True, 96, 3, 257), (True, False, True, 3, 96, 7),
])
def test_graph_linear_vs_oracle(cuda_device, typed, learn, bias, kin, kout, batch):
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200.testing import synth_state_dict
    spec = sdb.get_skeleton("amass")
    N = spec.num_nodes
    nt = spec.nodes_type_id if typed else None
    layer = sdb.StaticGraphLinear(kin, kout, bias=bias, num_nodes=N, node_types=nt, learn_influence=learn)
    sd = synth_state_dict(layer.state_dict(), seed=5, mode="perturbed", gain=1.5)
    layer.load_state_dict(sd)
    x = torch.randn(batch, N, kin, generator=torch.Generator().manual_seed(1))
    ref = oc.graph_linear({k: v for k, v in sd.items()}, "", x, nt, learn)
    out = layer.to(cuda_device)(x.to(cuda_device))
    assert G.rel_err(out.cpu(), ref) < FP32_TOL


def test_graph_linear_identity_path_and_epilogue(cuda_device):
    """G == I (fused epilogue, no scratch) with two K segments, in-place repeat, scale/shift, tanh, residual."""
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200 import _native as nv
    from skeletondiffusion_b200.testing import synth_state_dict
    spec = sdb.get_skeleton("h36m")
    N, nt = spec.num_nodes, spec.nodes_type_id
    layer = sdb.StaticGraphLinear(192, 192, bias=True, num_nodes=N, node_types=nt, learn_influence=True)
    sd = synth_state_dict(layer.state_dict(), seed=9, mode="perturbed", gain=1.5)
    sd["G"] = torch.eye(N)
    layer.load_state_dict(sd)
    g = torch.Generator().manual_seed(2)
    W, S = 3, 5
    cond = torch.randn(W, N, 96, generator=g)
    x = torch.randn(W * S, N, 96, generator=g)
    ss = torch.randn(4, 2 * 192, generator=g) * 0.3
    rows = torch.tensor([i % 4 for i in range(W * S)], dtype=torch.int32)
    res = torch.randn(W * S, N, 192, generator=g)
    xin = torch.cat([cond.repeat_interleave(S, 0), x], -1)
    y = oc.graph_linear(sd, "", xin, nt, True)
    scale, shift = ss[rows.long()][:, None, :192], ss[rows.long()][:, None, 192:]
    ref = torch.tanh(y * (scale + 1) + shift) + res
    plan = layer.to(cuda_device).plan()
    assert plan.identity
    d = cuda_device
    out = plan.forward(cond.to(d), x2=x.to(d), scale_shift=ss.to(d), ss_rows=rows.to(d), act=nv.ACT_TANH, residual=res.to(d), rep=S)
    assert G.rel_err(out.cpu(), ref) < FP32_TOL


@pytest.mark.parametrize("N,heads,dh", [(21, 8, 32), (16, 4, 32), (51, 8, 32), (17, 2, 16)])
def test_node_attention_vs_oracle(cuda_device, N, heads, dh):
    nv = _native()
    g = torch.Generator().manual_seed(3)
    B = 9
    qkv = torch.randn(B, N, 3 * heads * dh, generator=g)
    q, k, v = [t.reshape(B, N, heads, dh).permute(0, 2, 1, 3) for t in qkv.chunk(3, -1)]
    attn = torch.einsum("bhnc,bhjc->bhnj", q * dh ** -0.5, k).softmax(-1)
    ref = torch.einsum("bhnj,bhjd->bhnd", attn, v).permute(0, 2, 1, 3).reshape(B, N, heads * dh)
    qd = qkv.to(cuda_device)
    out = torch.empty(B, N, heads * dh, device=cuda_device)
    nv.check(nv.load().sd_node_attention(qd.data_ptr(), out.data_ptr(), B, N, heads, dh, nv.stream_ptr(cuda_device)), "attn")
    assert G.rel_err(out.cpu(), ref) < FP32_TOL


@pytest.mark.parametrize("N,B", [(21, 700), (16, 1001), (17, 149)])
def test_bulk_attention_ring_wraps(cuda_device, N, B):
    """The shipped configuration (8 heads x 32) runs the persistent bulk-copy kernel: one CTA per SM, samples grid-strided through
    a 3-stage ring.  Batches larger than the CTA count (and not multiples of it) make every stage be loaded, computed on, stored
    from and re-armed several times; every sample is checked against a plain fp32 softmax attention."""
    nv = _native()
    g = torch.Generator().manual_seed(N * 1000 + B)
    heads, dh = 8, 32
    qkv = torch.randn(B, N, 3 * heads * dh, generator=g)
    q, k, v = [t.reshape(B, N, heads, dh).permute(0, 2, 1, 3) for t in qkv.chunk(3, -1)]
    attn = torch.einsum("bhnc,bhjc->bhnj", q * dh ** -0.5, k).softmax(-1)
    ref = torch.einsum("bhnj,bhjd->bhnd", attn, v).permute(0, 2, 1, 3).reshape(B, N, heads * dh)
    qd = qkv.to(cuda_device)
    out = torch.full((B, N, heads * dh), float("nan"), device=cuda_device)
    nv.check(nv.load().sd_node_attention(qd.data_ptr(), out.data_ptr(), B, N, heads, dh, nv.stream_ptr(cuda_device)), "attn")
    assert torch.isfinite(out).all()                            # every row of every sample was written
    assert torch.equal(qd.cpu(), qkv)                           # the input is not modified (the in-place output lives in shared memory)
    per_sample = (out.cpu() - ref).abs().amax(dim=(1, 2)) / ref.abs().amax()
    assert float(per_sample.max()) < 1e-5


@pytest.mark.parametrize("N,iso", [(21, False), (16, False), (17, False), (19, False), (21, True)])
def test_reverse_step_vs_oracle(cuda_device, N, iso):
    """Templated (16/17/21), generic (19) and diagonal (U = I) variants; t = 0 must equal clamp(x0)."""
    import skeletondiffusion_b200 as sdb
    g = torch.Generator().manual_seed(4)
    rand = (torch.rand(N, N, generator=g) >= 0.5).float()
    corr = (rand + rand.T) // 2
    sigma, lam, u = sdb.get_cov_from_corr(corr, if_run_as_isotropic=iso)
    tab = oc.diffusion_tables(lam, u, 10)
    model = sdb.Denoiser(dim=96, cond_dim=0, out_dim=96, channels=N, num_nodes=N)
    diff = sdb.NonisotropicGaussianDiffusion(Sigma_N=sigma, Lambda_N=lam, U=u, model=model).to(cuda_device)
    B = 37
    x_t, x0, eps = (torch.randn(B, N, 96, generator=g) for _ in range(3))
    x0 = x0 * 1.5
    for t in (9, 4, 1, 0):
        ref, ref_mean = oc.reverse_step(tab, u, x_t, x0, t, eps if t > 0 else None)
        out, mean = diff._reverse_step(x_t.to(cuda_device), x0.to(cuda_device), eps.to(cuda_device) if t > 0 else None, t, True, want_mean=True)
        assert G.rel_err(out.cpu(), ref) < FP32_TOL and G.rel_err(mean.cpu(), ref_mean) < FP32_TOL
        if t == 0:
            assert G.rel_err(out.cpu(), x0.clamp(-1, 1)) < 1e-5


def test_fill_normal_statistics(cuda_device):
    nv = _native()
    n = 1 << 22
    a = torch.empty(n, device=cuda_device)
    b = torch.empty(n, device=cuda_device)
    nv.check(nv.load().sd_fill_normal(a.data_ptr(), n, 123, 0, nv.stream_ptr(cuda_device)), "fill")
    nv.check(nv.load().sd_fill_normal(b.data_ptr(), n // 2, 123, n // 8, nv.stream_ptr(cuda_device)), "fill")
    assert abs(float(a.mean())) < 3e-3 and abs(float(a.std()) - 1) < 3e-3
    assert abs(float((a ** 4).mean()) - 3.0) < 0.05
    assert torch.equal(a[n // 2:], b[:n // 2])          # counter-based: a shard reproduces its slice of the stream


# ------------------------------------------------------------------------------------------------
# composites against the reference's golden vectors
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", G.README_CASES)
def test_readme_sampling_golden(cuda_device, name):
    case = G.load_npz(name)
    diff, sd, _ = G.readme_models(case, device=cuda_device)
    d = cuda_device
    out = diff.model(case["x_probe"].to(d), case["t_probe"].to(d))
    assert G.rel_err(out.cpu(), case["den_out"]) < FP32_TOL
    lat, (n0, noise_t, mean_t) = diff.sample(batch_size=4, start_noise=case["start_noise"].to(d),
                                              sampling_noise=case["sampling_noise"].to(d), return_sampling_noise=True)
    assert torch.equal(n0.cpu(), case["start_noise"])
    assert G.rel_err(mean_t.cpu(), case["mean_t"]) < FP32_TOL
    assert G.rel_err(lat.cpu(), case["latents"]) < FP32_TOL


@pytest.mark.parametrize("name", G.DATASET_CASES)
def test_dataset_pipeline_golden(cuda_device, name):
    import skeletondiffusion_b200 as sdb
    case = G.load_npz(name)
    spec, ae, diff, ae_sd, diff_sd = G.dataset_models(case, device=cuda_device)
    d = cuda_device
    S, W, ph = int(case["samples"]), int(case["windows"]), int(case["ph"])
    obs = case["obs"].to(d)
    # (a9) encode
    z = ae.get_past_embedding(obs)
    assert G.rel_err(z.cpu(), case["z_past"]) < FP32_TOL
    # (a5) one Denoiser forward with per-sample times and in-place repeated conditioning
    out = diff.model(case["x_probe"].to(d), case["t_probe"].to(d), None, case["z_past"].to(d))
    assert G.rel_err(out.cpu(), case["den_out"]) < FP32_TOL
    # (a3/a4) whole reverse process with injected noise, per-step posterior means
    lat, (_, _, mean_t) = diff.sample(batch_size=W * S, x_cond=case["z_past"].to(d), start_noise=case["start_noise"].to(d),
                                      sampling_noise=case["sampling_noise"].to(d), return_sampling_noise=True)
    assert G.rel_err(mean_t.cpu(), case["mean_t"]) < FP32_TOL
    assert G.rel_err(lat.cpu(), case["latents"]) < FP32_TOL
    # (a10) decode from the reference's latents
    pred = ae.decode(obs, case["latents"].to(d), None, ph=ph)
    assert G.rel_err(pred.cpu().view(case["pred"].shape), case["pred"]) < FP32_TOL
    # (a12) the caller: get_prediction end to end, then ADE/FDE/APD to the printed precision (4 decimals, eval.py:109)
    pred2 = sdb.get_prediction(obs, (ae, diff), num_samples=S, pred_length=ph, diffusion_conditioning=True,
                               sampler_kwargs=dict(start_noise=case["start_noise"].to(d), sampling_noise=case["sampling_noise"].to(d)))
    assert tuple(pred2.shape) == tuple(case["pred"].shape)
    assert G.rel_err(pred2.cpu(), case["pred"]) < 2 * FP32_TOL
    pm, tm = spec.transform_to_metric_space(pred2.cpu()), spec.transform_to_metric_space(case["target"])
    for fn, key in ((lambda: oc.ade(tm, pm), "ade"), (lambda: oc.fde(tm, pm), "fde"), (lambda: oc.apd(pm), "apd")):
        assert torch.allclose(fn(), case[key], atol=5e-5, rtol=1e-4), key
    # (a11) training-loss entry point
    xq = diff.q_sample(case["x_start"].to(d), case["t_loss"].to(d), case["noise_loss"].to(d))
    assert G.rel_err(xq.cpu(), case["q_sample"]) < FP32_TOL
    loss, lw, mout = diff.p_losses(case["x_start"].to(d), case["t_loss"].to(d), noise=case["noise_loss"].to(d), x_cond=case["z_past"].to(d))
    assert G.rel_err(mout.cpu(), case["loss_model_out"]) < FP32_TOL
    assert G.rel_err(loss.cpu(), case["loss"]) < FP32_TOL
    assert torch.allclose(lw.cpu().reshape(-1), case["loss_weight"].reshape(-1))


# ------------------------------------------------------------------------------------------------
# edge cases and full-size, size-independent properties
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("batch", [1, 127, 129])
def test_ragged_batches_match_oracle(cuda_device, batch):
    case = G.load_npz("h36m_perturbed")
    spec, ae, diff, ae_sd, diff_sd = G.dataset_models(case, device=cuda_device)
    cfg = G.dataset_cfg(spec)
    g = torch.Generator().manual_seed(batch)
    x = torch.randn(batch, spec.num_nodes, 96, generator=g)
    cond = torch.tanh(torch.randn(batch, spec.num_nodes, 96, generator=g))
    t = torch.randint(0, 10, (batch,), generator=g)
    ref = oc.denoiser_forward(diff_sd, cfg, x[:3], t[:3], cond[:3], prefix="model.")
    out = diff.model(x.to(cuda_device), t.to(cuda_device), None, cond.to(cuda_device))
    assert G.rel_err(out[:3].cpu(), ref) < FP32_TOL
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
def test_empty_batch_returns_empty_tensors(cuda_device, precision):
    """Zero windows / zero samples: the reference's modules return empty tensors of the right shape; so do the kernels'
    wrappers (zero-size device buffers have null pointers: the C entry points treat batch == 0 as a no-op before any check)."""
    import skeletondiffusion_b200 as sdb
    spec = sdb.get_skeleton("h36m")
    ae, diff = sdb.build_models(spec, cuda_device, precision=precision)
    d, N = cuda_device, spec.num_nodes
    x = torch.zeros(0, N, 96, device=d)
    t = torch.zeros(0, dtype=torch.long, device=d)
    obs = torch.zeros(0, spec.obs_length, N, 3, device=d)
    assert tuple(diff.model(x, t, None, x, precision=precision).shape) == (0, N, 96)
    assert tuple(diff.model.layers[0][0].block2.proj.plan().forward(torch.zeros(0, N, 192, device=d), precision=precision).shape) == (0, N, 192)
    assert tuple(ae.get_past_embedding(obs).shape) == (0, N, 96)
    assert tuple(ae.decode(obs, x, None, ph=5).shape) == (0, 5, N, 3)
    assert tuple(diff.sample(batch_size=0, x_cond=x)[0].shape) == (0, N, 96)
    pred = sdb.get_prediction(obs, (ae, diff), num_samples=4, pred_length=5, diffusion_conditioning=True)
    assert tuple(pred.shape) == (0, 4, 5, N, 3)


def test_full_size_batch_independence(cuda_device):
    """BASELINE eval geometry (512 windows x 50 samples = 25 600 latents): every row of the big batch must equal
    the same row sampled in a small batch (no cross-sample coupling), the t=0 output is clamped, and
    the first rows still match the reference's golden latents."""
    case = G.load_npz("amass_perturbed")
    spec, ae, diff, _, _ = G.dataset_models(case, device=cuda_device)
    d = cuda_device
    W, S, N, T = 512, 50, spec.num_nodes, 10
    B = W * S
    g = torch.Generator(device="cpu").manual_seed(7)
    zp = torch.tanh(torch.randn(W, N, 96, generator=g)).to(d)
    zp[:2] = case["z_past"].to(d)
    start = torch.randn(B, N, 96, generator=g).to(d)
    noise = torch.randn(B, T - 1, N, 96, generator=g).to(d)
    # rows (window 0, samples 0..2) and (window 1, samples 0..2) replay the golden case
    gs, gn = case["start_noise"].to(d), case["sampling_noise"].to(d)
    idx = torch.tensor([0, 1, 2, S, S + 1, S + 2], device=d)
    start[idx], noise[idx] = gs, gn
    lat, _ = diff.sample(batch_size=B, x_cond=zp, start_noise=start, sampling_noise=noise)
    assert lat.shape == (B, N, 96) and torch.isfinite(lat).all()
    assert float(lat.abs().max()) <= 1.0 + 1e-5          # t=0: C1 = I up to 6e-7
    assert G.rel_err(lat[idx].cpu(), case["latents"]) < FP32_TOL
    rows = torch.tensor([5 * S + 3, 5 * S + 4, 300 * S + 49, 511 * S], device=d)     # windows 5, 5, 300, 511
    small, _ = diff.sample(batch_size=4, x_cond=zp[torch.tensor([5, 5, 300, 511], device=d)], start_noise=start[rows].clone(),
                           sampling_noise=noise[rows].clone())
    assert G.rel_err(lat[rows].cpu(), small.cpu()) < 1e-5


def test_step_kernel_linearity_full_size(cuda_device):
    """The step is linear in (x_t, eps) once x0 is fixed: step(a x + b y) = a step(x) + b step(y) at x0 = 0."""
    case = G.load_npz("amass_perturbed")
    spec, ae, diff, _, _ = G.dataset_models(case, device=cuda_device)
    B, N = 25600, spec.num_nodes
    g = torch.Generator(device=cuda_device).manual_seed(11)
    x, y, e1, e2 = (torch.randn(B, N, 96, device=cuda_device, generator=g) for _ in range(4))
    z = torch.zeros_like(x)
    s1, _ = diff._reverse_step(x, z, e1, 5)
    s2, _ = diff._reverse_step(y, z, e2, 5)
    s3, _ = diff._reverse_step(2.0 * x - 0.5 * y, z, 2.0 * e1 - 0.5 * e2, 5)
    assert G.rel_err(s3.cpu(), (2.0 * s1 - 0.5 * s2).cpu()) < 1e-5


@pytest.mark.parametrize("dataset", ["amass", "h36m"])
def test_autoencoder_identity_influence_fast_path(cuda_device, dataset):
    """Identity graph influence (the reference initialisation) selects the fused FFMA2 GRU step; the weights are still
    the dense per-type stress weights, so a wrong gate interleave or type index cannot hide."""
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200.testing import synth_state_dict
    spec = sdb.get_skeleton(dataset)
    ae, _ = sdb.build_models(spec, "cpu", seed=3)
    sd = synth_state_dict(ae.state_dict(), seed=77, mode="perturbed", gain=2.5)
    N = spec.num_nodes
    for k in sd:
        if k.endswith(".G"):
            sd[k] = torch.eye(N)
        if k.endswith(".G_add"):
            sd[k] = torch.zeros(N, N)
    ae.load_state_dict(sd)
    cfg = G.dataset_cfg(spec)
    g = torch.Generator().manual_seed(5)
    W, S, ph = 3, 4, 17
    obs = (torch.randn(W, spec.obs_length, N, 3, generator=g) * 0.3).clamp(-1, 1)
    lat = torch.tanh(torch.randn(W * S, N, 96, generator=g))
    z_ref = oc.encode(sd, cfg, obs)
    p_ref = oc.decode(sd, cfg, obs[:, -2:].repeat_interleave(S, 0), lat, ph)
    ae = ae.to(cuda_device).eval()
    assert ae.decoder.rnn.layers[0].plan(ph).identity
    z = ae.get_past_embedding(obs.to(cuda_device))
    p = ae.decode(obs.to(cuda_device), lat.to(cuda_device), None, ph=ph)
    assert G.rel_err(z.cpu(), z_ref) < FP32_TOL
    assert G.rel_err(p.cpu(), p_ref) < FP32_TOL


def test_cpu_tensor_is_rejected(cuda_device):
    import skeletondiffusion_b200 as sdb
    nv = _native()
    layer = sdb.StaticGraphLinear(8, 8, num_nodes=4).to(cuda_device)
    with pytest.raises(nv.NativeError):
        layer(torch.zeros(2, 4, 8))


def test_cuda_graph_sampling_matches_eager(cuda_device):
    """The whole p_sample_loop captured as one CUDA graph (replayed twice with different inputs) equals the eager loop."""
    case = G.load_npz("h36m_perturbed")
    spec, ae, diff, _, _ = G.dataset_models(case, device=cuda_device)
    d = cuda_device
    W, S = int(case["windows"]), int(case["samples"])
    kw = dict(batch_size=W * S, x_cond=case["z_past"].to(d), start_noise=case["start_noise"].to(d), sampling_noise=case["sampling_noise"].to(d))
    eager, _ = diff.sample(**kw)
    diff.use_cuda_graph = True
    g1, _ = diff.sample(**kw)
    kw2 = dict(kw, start_noise=kw["start_noise"].flip(0).contiguous())
    diff.use_cuda_graph = False
    eager2, _ = diff.sample(**kw2)
    diff.use_cuda_graph = True
    g2, _ = diff.sample(**kw2)
    assert len(diff._graphs) == 1
    assert torch.equal(g1, eager) and torch.equal(g2, eager2)
    assert G.rel_err(g1.cpu(), case["latents"]) < FP32_TOL


def test_noise_does_not_depend_on_the_sharding(cuda_device):
    """Philox noise is indexed by the GLOBAL row: a shard that passes its row / window offset draws exactly what the unsharded call
    draws for those rows (results independent of the rank count), for sample() and for the captured pipeline graph."""
    import skeletondiffusion_b200 as sdb
    d = cuda_device
    spec = sdb.get_skeleton("h36m")
    ae, diff = sdb.build_models(spec, d, precision="fp16x2")
    N, S, W = spec.num_nodes, 5, 6
    cond = torch.tanh(torch.randn(W, N, 96, device=d))
    torch.manual_seed(11); diff._noise_calls = 0
    full, _ = diff.sample(batch_size=W * S, x_cond=cond)
    torch.manual_seed(11); diff._noise_calls = 0
    part, _ = diff.sample(batch_size=2 * S, x_cond=cond[4:6].contiguous(), noise_row_offset=4 * S)
    assert torch.equal(part, full[4 * S:])
    obs = (torch.randn(W, spec.obs_length, N, 3, device=d) * 0.3).clamp(-1, 1)
    g_full, g_part = sdb.GraphedPrediction((ae, diff), W, S, 4, d), sdb.GraphedPrediction((ae, diff), 2, S, 4, d)
    torch.manual_seed(5); diff._noise_calls = 0
    p_full = g_full(obs, clone=True)
    torch.manual_seed(5); diff._noise_calls = 0
    p_part = g_part(obs[2:4].contiguous(), clone=True, window_offset=2)
    assert torch.equal(p_part, p_full[2:4])


def test_invalidate_plans_after_a_data_write(cuda_device):
    """A `.data` write does not bump autograd's version counter, so the packed copies go stale until invalidate_plans()."""
    import skeletondiffusion_b200 as sdb
    d = cuda_device
    spec = sdb.get_skeleton("h36m")
    layer = sdb.StaticGraphLinear(96, 96, bias=True, num_nodes=spec.num_nodes, node_types=spec.nodes_type_id, learn_influence=True).to(d)
    x = torch.randn(9, spec.num_nodes, 96, device=d)
    a = layer(x, precision="fp16x2")
    layer.weight.data.mul_(2.0)
    sdb.invalidate_plans()
    b = layer(x, precision="fp16x2")
    assert not torch.allclose(a, b)
    ref = oc.graph_linear({k: v.detach().cpu() for k, v in layer.state_dict().items()}, "", x.cpu(), spec.nodes_type_id, True)
    assert G.rel_err(b.cpu(), ref) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "fp16x2"])
def test_amass_mano_reachability_pipeline_vs_oracle(cuda_device, precision):
    """AMASS-MANO (configs/config_eval/dataset/amass-mano.yaml: 51 nodes, 43 node types) with covariance_matrix_type='reachability'
    (kinematic/base.py:85-127), dense graph influence and typed weights: the whole pipeline against the oracle on the same injected
    noise.  N = 51 has no specialised per-sample kernels: this is the generic path of every layer."""
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200.testing import synth_state_dict
    d = cuda_device
    spec = sdb.get_skeleton("amass-mano")
    N, nt = spec.num_nodes, spec.nodes_type_id
    torch.manual_seed(3)
    ae = sdb.AutoEncoder(num_nodes=N, encoder_hidden_size=96, decoder_hidden_size=96, latent_size=96, node_types=nt, input_size=3,
                         z_activation="tanh", enc_num_layers=spec.enc_num_layers, recurrent_arch_enc="StaticGraphGRU",
                         recurrent_arch_decoder="StaticGraphGRU", output_size=3, if_consider_hip=False)
    mgr = sdb.DiffusionManager(diffusion_type="NonisotropicGaussianDiffusion", skeleton=spec, covariance_matrix_type="reachability",
                               reachability_matrix_degree_factor=0.5, reachability_matrix_stop_at="hips", num_nodes=N, node_types=nt,
                               diffusion_conditioning=True, latent_size=96, diffusion_timesteps=10, diffusion_objective="pred_x0",
                               beta_schedule="cosine", precision=precision,
                               diffusion_arch=dict(depth=2, attn_heads=8, attn_dim_head=32, use_attention=True, self_condition=False,
                                                   norm_type="none", learn_influence=True))
    diff = mgr.get_diffusion()
    tabs = {k: v.clone() for k, v in diff.state_dict().items() if not k.startswith("model.")}
    diff_sd = synth_state_dict(diff.state_dict(), seed=61, mode="perturbed", gain=1.5)
    diff_sd.update(tabs)
    ae_sd = synth_state_dict(ae.state_dict(), seed=62, mode="perturbed", gain=1.5)
    diff.load_state_dict(diff_sd); ae.load_state_dict(ae_sd)
    W, S, ph = 3, 4, 5
    g = torch.Generator().manual_seed(9)
    obs = (torch.randn(W, spec.obs_length, N, 3, generator=g) * 0.3).clamp(-1, 1)
    start = torch.randn(W * S, N, 96, generator=g)
    noise = torch.randn(W * S, 9, N, 96, generator=g)
    cfg = dict(dim=96, cond_dim=96, depth=2, attn_heads=8, attn_dim_head=32, node_types=nt, learn_influence=True, enc_num_layers=spec.enc_num_layers)
    ref = oc.get_prediction(ae_sd, diff_sd, cfg, tabs, tabs["U"], obs, S, ph, start, noise)
    ae, diff = ae.to(d).eval(), diff.to(d).eval()
    pred = sdb.get_prediction(obs.to(d), (ae, diff), num_samples=S, pred_length=ph, diffusion_conditioning=True,
                              sampler_kwargs=dict(start_noise=start.to(d), sampling_noise=noise.to(d)))
    assert G.rel_err(pred.cpu(), ref) < FP32_TOL
