"""GPU parity of the general (dense) graph-influence path: per-sample mix kernel, mix fused into the attention kernel,
per-sample GRU gate kernel + output head.  Through the C ABI, against the CPU oracle / float64 torch on the same inputs.
Reference semantics: graph_structural.py:30-43 (mix after the per-type products), recurrent.py:333-363 (GRU cell with gx_i)."""
import pytest
import torch

from oracle import skeldiff_oracle as oc
from tests import _golden as G

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4


def _layer(spec, kin, kout, bias, seed):
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200.testing import synth_state_dict
    layer = sdb.StaticGraphLinear(kin, kout, bias=bias, num_nodes=spec.num_nodes, node_types=spec.nodes_type_id, learn_influence=True)
    sd = synth_state_dict(layer.state_dict(), seed=seed, mode="perturbed", gain=1.5)
    layer.load_state_dict(sd)
    return layer, sd


@pytest.mark.parametrize("dataset", ["amass", "h36m", "freeman"])
@pytest.mark.parametrize("kout,precision", [(192, "fp32"), (192, "bf16x3"), (96, "bf16x3"), (288, "fp32"), (192, "fp16x2"), (288, "fp16x2")])
def test_dense_graph_linear_epilogue_vs_oracle(cuda_device, dataset, kout, precision):
    """Dense G^: raw products + sample_mix_kernel with bias, scale/shift, tanh and residual (64- and 32-column tasks)."""
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200 import _native as nv
    spec = sdb.get_skeleton(dataset)
    N, nt = spec.num_nodes, spec.nodes_type_id
    layer, sd = _layer(spec, 192, kout, True, 21)
    g = torch.Generator().manual_seed(3)
    B = 301                                         # not a multiple of the grid: CTAs own different sample counts
    x = torch.randn(B, N, 192, generator=g)
    ss = torch.randn(1, 2 * kout, generator=g) * 0.3
    res = torch.randn(B, N, kout, generator=g)
    y = oc.graph_linear(sd, "", x, nt, True)
    ref = torch.tanh(y * (ss[0, :kout] + 1) + ss[0, kout:]) + res
    plan = layer.to(cuda_device).plan()
    assert not plan.identity
    d = cuda_device
    out = plan.forward(x.to(d), scale_shift=ss.to(d), act=nv.ACT_TANH, residual=res.to(d), precision=precision)
    assert G.rel_err(out.cpu(), ref) < FP32_TOL
    # no activation / no residual variant, and the in-place residual the Denoiser uses (out aliases the residual)
    out2 = plan.forward(x.to(d), precision=precision)
    assert G.rel_err(out2.cpu(), y) < FP32_TOL
    buf = res.to(d).clone()
    plan.forward(x.to(d), residual=buf, out=buf, precision=precision)
    assert G.rel_err(buf.cpu(), y + res) < FP32_TOL


def test_dense_mix_is_repeatable_and_batch_independent_at_full_size(cuda_device):
    """B = 25 600 (the AMASS eval batch): 20 launches give bitwise identical results (a stage released too early or a lost
    barrier phase shows up as a run-to-run difference), and a sample's rows do not depend on its neighbours."""
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200 import _native as nv
    spec = sdb.get_skeleton("amass")
    N = spec.num_nodes
    layer, _ = _layer(spec, 192, 192, True, 4)
    plan = layer.to(cuda_device).plan()
    d = cuda_device
    B = 25600
    x = torch.randn(B, N, 192, device=d)
    res = torch.randn(B, N, 192, device=d)
    ss = torch.randn(1, 384, device=d) * 0.3
    first = plan.forward(x, scale_shift=ss, act=nv.ACT_TANH, residual=res, precision="bf16x3").clone()
    for _ in range(20):
        again = plan.forward(x, scale_shift=ss, act=nv.ACT_TANH, residual=res, precision="bf16x3")
        assert torch.equal(again, first)
    idx = torch.tensor([0, 147, 148, 12345, 25599], device=d)
    sub = plan.forward(x[idx].contiguous(), scale_shift=ss, act=nv.ACT_TANH, residual=res[idx].contiguous(), precision="bf16x3")
    assert torch.equal(sub, first[idx])


@pytest.mark.parametrize("dataset", ["amass", "h36m", "freeman"])
def test_attention_with_fused_qkv_mix_vs_oracle(cuda_device, dataset):
    """Residual(PreNorm(Attention)) with dense G^ on to_qkv / to_out: the to_qkv mix runs inside the attention kernel."""
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200.testing import synth_state_dict
    spec = sdb.get_skeleton(dataset)
    N, nt = spec.num_nodes, spec.nodes_type_id
    att = sdb.network.Residual(sdb.network.PreNorm(192, sdb.network.Attention(192, heads=8, dim_head=32, num_nodes=N, node_types=nt, learn_influence=True)))
    sd = synth_state_dict(att.state_dict(), seed=11, mode="perturbed", gain=1.5)
    att.load_state_dict(sd)
    x = torch.randn(77, N, 192, generator=torch.Generator().manual_seed(8))
    ref = oc._node_attention(sd, "fn.", x, 8, 32, nt, True)          # includes the residual (attention.py:16-17)
    for precision in ("fp32", "bf16x3", "fp16x2"):
        out = att.to(cuda_device)(x.to(cuda_device), precision=precision)
        assert G.rel_err(out.cpu(), ref) < FP32_TOL, precision


def test_denoiser_dense_influence_is_repeatable_at_full_size(cuda_device):
    """The whole Denoiser with dense influence matrices at B = 25 600: two forwards are bitwise identical."""
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200.testing import synth_state_dict
    spec = sdb.get_skeleton("amass")
    _, diff = sdb.build_models(spec, "cpu")
    diff.load_state_dict(synth_state_dict(diff.state_dict(), seed=1, mode="perturbed", gain=2.5))
    model = diff.model.to(cuda_device).eval()
    d = cuda_device
    x = torch.randn(25600, spec.num_nodes, 96, device=d)
    cond = torch.tanh(torch.randn(512, spec.num_nodes, 96, device=d))
    t = torch.full((25600,), 5, device=d, dtype=torch.long)
    plan = model.plan()
    a = plan.forward(x, cond, 5, precision="bf16x3").clone()
    for _ in range(5):
        b = plan.forward(x, cond, 5, precision="bf16x3")
        assert torch.equal(a, b)
    del t


@pytest.mark.parametrize("dataset", ["amass", "h36m", "freeman"])
@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "fp16x2"])
def test_decode_dense_influence_vs_oracle(cuda_device, dataset, precision):
    """Decoder with dense G / G_add / fc.G (gx_i changes every frame): tcgen05 recurrent product (bf16x3) or FFMA (fp32),
    gru_sample_kernel, gru_head_kernel; 25 frames so that the recurrence accumulates."""
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200.testing import synth_state_dict
    spec = sdb.get_skeleton(dataset)
    ae, _ = sdb.build_models(spec, "cpu")
    sd = synth_state_dict(ae.state_dict(), seed=6, mode="perturbed", gain=2.5)
    ae.load_state_dict(sd)
    g = torch.Generator().manual_seed(4)
    W, S, ph = 3, 7, 25
    obs = (torch.randn(W, spec.obs_length, spec.num_nodes, 3, generator=g) * 0.3).clamp(-1, 1)
    lat = torch.tanh(torch.randn(W * S, spec.num_nodes, 96, generator=g))
    cfg = G.dataset_cfg(spec)
    ref = oc.decode(sd, cfg, obs.repeat_interleave(S, 0), lat, ph)
    out = ae.to(cuda_device).decode(obs.to(cuda_device), lat.to(cuda_device), None, ph=ph, precision=precision)
    assert G.rel_err(out.cpu(), ref) < FP32_TOL
    # identity influence on the same kernels (tensor-core precisions take the per-sample path for every cell)
    ae2, _ = sdb.build_models(spec, "cpu")
    sd2 = synth_state_dict(ae2.state_dict(), seed=6, mode="perturbed", gain=2.5)
    for k in list(sd2):
        if k.endswith(".G"):
            sd2[k] = torch.eye(spec.num_nodes)
        if k.endswith(".G_add"):
            sd2[k] = torch.zeros(spec.num_nodes, spec.num_nodes)
    ae2.load_state_dict(sd2)
    ref2 = oc.decode(sd2, cfg, obs.repeat_interleave(S, 0), lat, ph)
    out2 = ae2.to(cuda_device).decode(obs.to(cuda_device), lat.to(cuda_device), None, ph=ph, precision=precision)
    assert G.rel_err(out2.cpu(), ref2) < FP32_TOL


@pytest.mark.parametrize("dataset", ["amass", "h36m", "freeman"])
@pytest.mark.parametrize("dense_head", [False, True])
def test_decode_fused_tensor_core_step_vs_oracle(cuda_device, dataset, dense_head):
    """Identity influence on the GRU cell under fp16x2: ONE tcgen05 kernel per frame (recurrent product with gate-interleaved
    weight rows + gates in the epilogue, operands through the residual ring, new state by bulk tensor store).  407 rows =
    three full 128-sample tiles and a ragged one per node; typed weights differ per node; the output head's own influence is
    identity or dense.  Against the CPU oracle (recurrent.py:333-358) and, tighter, against the exact-fp32 path."""
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200.testing import synth_state_dict
    spec = sdb.get_skeleton(dataset)
    ae, _ = sdb.build_models(spec, "cpu")
    sd = synth_state_dict(ae.state_dict(), seed=11, mode="perturbed", gain=2.5)
    for k in list(sd):
        if ".rnn." in k and k.endswith(".G"):
            sd[k] = torch.eye(spec.num_nodes)
        if ".rnn." in k and k.endswith(".G_add"):
            sd[k] = torch.zeros(spec.num_nodes, spec.num_nodes)
        if not dense_head and k.endswith(".G") and ".rnn." not in k:
            sd[k] = torch.eye(spec.num_nodes)
    ae.load_state_dict(sd)
    g = torch.Generator().manual_seed(9)
    W, S, ph = 11, 37, 12
    obs = (torch.randn(W, spec.obs_length, spec.num_nodes, 3, generator=g) * 0.3).clamp(-1, 1)
    lat = torch.tanh(torch.randn(W * S, spec.num_nodes, 96, generator=g))
    ref = oc.decode(sd, G.dataset_cfg(spec), obs.repeat_interleave(S, 0), lat, ph)
    ae = ae.to(cuda_device)
    out = ae.decode(obs.to(cuda_device), lat.to(cuda_device), None, ph=ph, precision="fp16x2")
    exact = ae.decode(obs.to(cuda_device), lat.to(cuda_device), None, ph=ph, precision="fp32")
    assert torch.isfinite(out).all()
    assert G.rel_err(out.cpu(), ref) < FP32_TOL
    assert G.rel_err(out.cpu(), exact.cpu()) < 2e-5
    again = ae.decode(obs.to(cuda_device), lat.to(cuda_device), None, ph=ph, precision="fp16x2")
    assert torch.equal(out, again)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "fp16x2"])
def test_encode_dense_influence_vs_oracle(cuda_device, precision):
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200.testing import synth_state_dict
    spec = sdb.get_skeleton("amass")
    ae, _ = sdb.build_models(spec, "cpu")
    sd = synth_state_dict(ae.state_dict(), seed=16, mode="perturbed", gain=2.5)
    ae.load_state_dict(sd)
    obs = (torch.randn(37, spec.obs_length, spec.num_nodes, 3, generator=torch.Generator().manual_seed(5)) * 0.3).clamp(-1, 1)
    ref = oc.encode(sd, G.dataset_cfg(spec), obs)
    out = ae.to(cuda_device).get_past_embedding(obs.to(cuda_device), precision=precision)
    assert G.rel_err(out.cpu(), ref) < FP32_TOL


def test_second_device_in_one_process(cuda_device):
    """the shared-memory opt-in is a per-device attribute and launches go to the tensors' device, not the current one."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import skeletondiffusion_b200 as sdb
    spec = sdb.get_skeleton("h36m")
    layer, sd = _layer(spec, 192, 192, True, 2)
    x = torch.randn(200, spec.num_nodes, 192, generator=torch.Generator().manual_seed(1))
    ref = oc.graph_linear(sd, "", x, spec.nodes_type_id, True)
    import copy
    for dev in ("cuda:0", "cuda:1"):
        l2 = copy.deepcopy(layer).to(dev)
        with torch.cuda.device(0):                  # current device stays 0 while the tensors live on `dev`
            out = l2.plan().forward(x.to(dev), precision="bf16x3")
        assert G.rel_err(out.cpu(), ref) < FP32_TOL, dev
