"""Few launches of the general graph-influence kernels at B = 25 600 for an ncu --set full capture."""
import sys, torch
sys.path.insert(0, '.')
import skeletondiffusion_b200 as sdb
from skeletondiffusion_b200 import _native as nv
from skeletondiffusion_b200.testing import synth_state_dict
dev = torch.device("cuda:0")
import os
PREC = os.environ.get("SKELDIFF_PRECISION", "fp16x2")
spec = sdb.get_skeleton('amass')
N, nt, B = spec.num_nodes, spec.nodes_type_id, 25600
l = sdb.StaticGraphLinear(192, 192, bias=True, num_nodes=N, node_types=nt, learn_influence=True)
l.load_state_dict(synth_state_dict(l.state_dict(), seed=1, mode="perturbed", gain=1.0))
p = l.to(dev).plan()
x = torch.randn(B, N, 192, device=dev); r = torch.randn(B, N, 192, device=dev); o = torch.empty(B, N, 192, device=dev)
ss = torch.randn(1, 384, device=dev) * 0.3
for _ in range(int(os.environ.get("PROF_ITERS", "2"))):
    p.forward(x, scale_shift=ss, act=nv.ACT_TANH, out=o, precision=PREC)
    p.forward(x, act=nv.ACT_TANH, residual=r, out=o, precision=PREC)
att = sdb.network.Residual(sdb.network.PreNorm(192, sdb.network.Attention(192, heads=8, dim_head=32, num_nodes=N, node_types=nt, learn_influence=True)))
att.load_state_dict(synth_state_dict(att.state_dict(), seed=3, mode="perturbed", gain=1.0))
att = att.to(dev)
for _ in range(int(os.environ.get("PROF_ITERS", "2"))):
    att(x, precision=PREC)
del x, r, o
ae, _ = sdb.build_models(spec, "cpu")
ae.load_state_dict(synth_state_dict(ae.state_dict(), seed=2, mode="perturbed", gain=2.5))
ae = ae.to(dev).eval()
obs = (torch.randn(512, spec.obs_length, N, 3, device=dev) * 0.3).clamp(-1, 1)
lat = torch.tanh(torch.randn(B, N, 96, device=dev))
ae.decode(obs, lat, None, ph=3, precision=PREC)
torch.cuda.synchronize()
print("ok")
if os.environ.get("PROF_DENOISER", "0") == "1":      # one dense Denoiser forward (the fused qkv-mix attention lives in sd_denoiser_forward)
    _, diff = sdb.build_models(spec, "cpu")
    diff.load_state_dict(synth_state_dict(diff.state_dict(), seed=1, mode="perturbed", gain=2.5))
    model = diff.model.to(dev).eval()
    xx = torch.randn(B, N, 96, device=dev)
    cond = torch.tanh(torch.randn(512, N, 96, device=dev))
    model.plan().forward(xx, cond, 5, precision=PREC)
    torch.cuda.synchronize()
    print("denoiser ok")
