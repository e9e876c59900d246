"""CUDA-event times of the general graph-influence kernels at the AMASS eval batch (B = 25 600): graph-linear pairs
(tc3 raw product + sample_mix) against the identity-influence fused kernel, attention with and without the fused qkv mix,
and one decoder step (recurrent product + gate kernel + head).  python scratch/time_mix.py > profiles/<name>.json"""
import json, sys, torch
sys.path.insert(0, '.')
import skeletondiffusion_b200 as sdb
from skeletondiffusion_b200 import _native as nv
from skeletondiffusion_b200.testing import synth_state_dict
from bench import _timed_kernel
dev = torch.device('cuda:0')
spec = sdb.get_skeleton('amass')
N, nt, B = spec.num_nodes, spec.nodes_type_id, 25600
res = {}
g = torch.Generator().manual_seed(0)


def layer(kin, kout, dense, bias=True):
    l = sdb.StaticGraphLinear(kin, kout, bias=bias, num_nodes=N, node_types=nt, learn_influence=True)
    sd = synth_state_dict(l.state_dict(), seed=kin + kout, mode="perturbed", gain=1.0)
    if not dense:
        sd["G"] = torch.eye(N)
    l.load_state_dict(sd)
    return l.to(dev).plan()


x = torch.randn(B, N, 192, device=dev)
r = torch.randn(B, N, 192, device=dev)
o = torch.empty(B, N, 192, device=dev)
ss = torch.randn(1, 384, device=dev) * 0.3
for dense in (False, True):
    p = layer(192, 192, dense)
    tag = "dense" if dense else "identity"
    res[f"glin192_{tag}_bare"] = _timed_kernel(dev, lambda: p.forward(x, out=o, precision="bf16x3")) * 1e3
    res[f"glin192_{tag}_ss_tanh"] = _timed_kernel(dev, lambda: p.forward(x, scale_shift=ss, act=nv.ACT_TANH, out=o, precision="bf16x3")) * 1e3
    res[f"glin192_{tag}_tanh_res"] = _timed_kernel(dev, lambda: p.forward(x, act=nv.ACT_TANH, residual=r, out=o, precision="bf16x3")) * 1e3
    p2 = layer(256, 192, dense, bias=False)
    a = torch.randn(B, N, 256, device=dev)
    res[f"to_out_{tag}_res_inplace"] = _timed_kernel(dev, lambda: p2.forward(a, residual=o, out=o, precision="bf16x3")) * 1e3
    del a
# attention block: to_qkv + attention (+ fused mix)
for dense in (False, True):
    tag = "dense" if dense else "identity"
    att = sdb.network.Residual(sdb.network.PreNorm(192, sdb.network.Attention(192, heads=8, dim_head=32, num_nodes=N, node_types=nt, learn_influence=True)))
    sd = synth_state_dict(att.state_dict(), seed=3, mode="perturbed", gain=1.0)
    if not dense:
        for k in sd:
            if k.endswith(".G"):
                sd[k] = torch.eye(N)
    att.load_state_dict(sd)
    att = att.to(dev)
    res[f"attention_block_{tag}"] = _timed_kernel(dev, lambda: att(x, precision="bf16x3"), iters=5) * 1e3
del x, r, o
torch.cuda.empty_cache()
# one decoder step = decode(ph) / ph
for dense in (False, True):
    tag = "dense" if dense else "identity"
    ae, _ = sdb.build_models(spec, "cpu")
    if dense:
        ae.load_state_dict(synth_state_dict(ae.state_dict(), seed=2, mode="perturbed", gain=2.5))
    ae = ae.to(dev).eval()
    obs = (torch.randn(512, spec.obs_length, N, 3, device=dev) * 0.3).clamp(-1, 1)
    lat = torch.tanh(torch.randn(B, N, 96, device=dev))
    for prec in ("bf16x3", "fp32"):
        if dense and prec == "fp32":
            continue
        ph = 20
        res[f"decode_step_{tag}_{prec}"] = _timed_kernel(dev, lambda: ae.decode(obs, lat, None, ph=ph, precision=prec), iters=3) * 1e3 / ph
print(json.dumps({k: round(v, 4) for k, v in res.items()}, indent=1))
