"""Signed error of one StaticGraphLinear (K = 192) on the FFMA and bf16x3 paths against a float64 product.

Separates random rounding (zero-mean, grows like sqrt(K)) from a systematic rounding direction in the tensor core's
fp32 accumulation (non-zero mean error of |y|, grows like the number of accumulate steps)."""
import sys, torch
sys.path.insert(0, '.')
import skeletondiffusion_b200 as sdb

dev = torch.device('cuda:0')
spec = sdb.get_skeleton('amass')
ae, diff = sdb.build_models(spec, dev)
plan = diff.model.layers[0][0].block2.proj.plan()
B, N, K = 8192, spec.num_nodes, plan.in_features
torch.manual_seed(0)
x = torch.randn(B, N, K, device=dev)
types = torch.tensor(list(plan.types_host), device=dev)
W = plan.weight.double()[types]                                   # [N, out, in]
y = torch.einsum('bmk,mok->bmo', x.double(), W)
if plan.bias_node is not None and plan.identity:
    y = y + plan.bias_node.double()
if not plan.identity:
    y = torch.einsum('nm,bmo->bno', plan.g.double(), y) + (plan.bias_node.double() if plan.bias_node is not None else 0)
for prec in ("fp32", "bf16x3"):
    o = plan.forward(x, precision=prec).double()
    e = o - y
    ulp = torch.finfo(torch.float32).eps * y.abs().clamp_min(1e-30)
    toward_zero = (e * torch.sign(y)) / ulp                       # < 0: magnitude lost (truncation)
    print(f"{prec:7s} rms err/|y| {float((e.norm() / y.norm())):.3e}   mean signed err (ulp of y, + = away from zero) "
          f"{float(toward_zero.median()):+.3f}   rms (ulp) {float((e / ulp).pow(2).mean().sqrt()):.2f}")
