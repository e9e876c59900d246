"""bf16x3 (three bf16 planes, 6 MMAs) against fp16x2 (two fp16 planes, 3 MMAs): single layers and the pipeline phases."""
import json, sys, torch
sys.path.insert(0, '.')
import skeletondiffusion_b200 as sdb
from skeletondiffusion_b200 import _native as nv
from skeletondiffusion_b200.testing import synth_state_dict
import bench
dev = torch.device('cuda:0')
spec = sdb.get_skeleton('amass')
N, nt = spec.num_nodes, spec.nodes_type_id
B = 25600
out = {}
g = torch.Generator().manual_seed(0)
for name, kin, kout, kw in (("glin192_bare", 192, 192, {}), ("glin192_tanh_res", 192, 192, dict(act=nv.ACT_TANH, res=True)),
                            ("to_qkv_192_768", 192, 768, {}), ("to_out_256_192_res", 256, 192, dict(res=True)), ("gru_96_288", 96, 288, {})):
    layer = sdb.StaticGraphLinear(kin, kout, bias=False, num_nodes=N, node_types=nt, learn_influence=True)
    sd = synth_state_dict(layer.state_dict(), seed=3, mode="perturbed", gain=1.0)
    sd["G"] = torch.eye(N)
    layer.load_state_dict(sd)
    plan = layer.to(dev).plan()
    x = torch.randn(B, N, kin, device=dev)
    res = torch.randn(B, N, kout, device=dev) if kw.get("res") else None
    o = torch.empty(B, N, kout, device=dev)
    for prec in ("bf16x3", "fp16x2"):
        t = bench._timed_kernel(dev, lambda: plan.forward(x, act=kw.get("act", nv.ACT_NONE), residual=res, out=o, precision=prec))
        out[f"{name}_{prec}"] = round(t * 1e3, 4)
    del x, res, o
print(json.dumps(out, indent=1))
