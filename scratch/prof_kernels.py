"""Profiling driver: runs selected kernels a few times at the full batch (for ncu captures) and prints CUDA-event times."""
import sys, torch
sys.path.insert(0, '.')
import skeletondiffusion_b200 as sdb
from skeletondiffusion_b200 import _native as nv
dev = torch.device('cuda:0')
spec = sdb.get_skeleton('amass')
ae, diff = sdb.build_models(spec, dev)
lib = nv.load()
B, N, C = 25600, spec.num_nodes, 192
layer = diff.model.layers[0][0].block2.proj
plan = layer.plan()
x = torch.randn(B, N, C, device=dev)
res = torch.randn(B, N, C, device=dev)
out = torch.empty(B, N, C, device=dev)
x16, r16 = x.to(torch.bfloat16), res.to(torch.bfloat16)
o16 = torch.empty_like(x16)
ss = torch.zeros(1, 2 * C, device=dev)
st = nv.stream_ptr(dev)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
import os
PREC = os.environ.get("SKELDIFF_PRECISION", "fp16x2")
x_t, x0, eps = (torch.randn(B, N, 96, device=dev) for _ in range(3))
qkv_t = torch.randn(B, N, 768, device=dev)
att = torch.empty(B, N, 256, device=dev)


def tc():
    nv.check(lib.sd_glin_forward_bf16(plan.handle, x16.data_ptr(), None, ss.data_ptr(), nv.ACT_TANH, r16.data_ptr(), o16.data_ptr(), 0, None, B, st), "tc")


def tc3():
    plan.forward(x, scale_shift=ss, act=nv.ACT_TANH, residual=res, out=out, precision=PREC)


def tc3nr():
    plan.forward(x, scale_shift=ss, act=nv.ACT_TANH, out=out, precision=PREC)


def tc3raw():
    plan.forward(x, out=out, precision=PREC)


_qkv = {}


def qkv():
    """to_qkv shape: 192 -> 768, row scale only"""
    if not _qkv:
        att = diff.model.layers[0][1]
        mod = att.fn.fn if hasattr(att, "fn") and hasattr(att.fn, "fn") else att
        while not hasattr(mod, "to_qkv"):
            mod = mod.fn
        _qkv["plan"] = mod.to_qkv.plan()
        _qkv["out"] = torch.empty(B, N, 768, device=dev)
        _qkv["rs"] = torch.rand(B, N, device=dev) + 0.5
    _qkv["plan"].forward(x, row_scale=_qkv["rs"], out=_qkv["out"], precision=PREC)


_to = {}


def toout():
    """to_out shape: 256 -> 192 + residual"""
    if not _to:
        att = diff.model.layers[0][1]
        mod = att
        while not hasattr(mod, "to_out"):
            mod = mod.fn
        _to["plan"] = mod.to_out.plan()
        _to["x"] = torch.randn(B, N, 256, device=dev)
    _to["plan"].forward(_to["x"], residual=res, out=out, precision=PREC)


def ffma():
    plan.forward(x, scale_shift=ss, act=nv.ACT_TANH, residual=res, out=out, precision="fp32")


def step():
    diff._reverse_step(x_t, x0, eps, 5)


def attn():
    nv.check(lib.sd_node_attention(qkv_t.data_ptr(), att.data_ptr(), B, N, 8, 32, st), "attn")


fns = {"toout": toout, "qkv": qkv, "tc": tc, "tc3": tc3, "tc3nr": tc3nr, "tc3raw": tc3raw, "ffma": ffma, "step": step, "attn": attn}
sel = list(fns) if which == "all" else which.split(",")
for name in sel:
    fn = fns[name]
    for _ in range(int(os.environ.get("PROF_WARM", "2"))):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = int(os.environ.get("PROF_ITERS", "5"))
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / iters * 1e3:.1f} us")
