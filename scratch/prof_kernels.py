"""Profiling driver: runs the dominant kernels a few times at the full batch (for ncu captures)."""
import sys, torch
sys.path.insert(0, '.')
import skeletondiffusion_b200 as sdb
from skeletondiffusion_b200 import _native as nv
dev = torch.device('cuda:0')
spec = sdb.get_skeleton('amass')
ae, diff = sdb.build_models(spec, dev)
lib = nv.load()
B, N, C = 25600, spec.num_nodes, 192
plan = diff.model.layers[0][0].block2.proj.plan()
x16 = torch.randn(B, N, C, device=dev).to(torch.bfloat16)
r16 = torch.randn(B, N, C, device=dev).to(torch.bfloat16)
o16 = torch.empty_like(x16)
ss = torch.zeros(2 * C, device=dev)
st = nv.stream_ptr(dev)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
x_t, x0, eps = (torch.randn(B, N, 96, device=dev) for _ in range(3))
for _ in range(4):
    if which in ("all", "tc"):
        nv.check(lib.sd_glin_forward_bf16(plan.handle, x16.data_ptr(), None, ss.data_ptr(), nv.ACT_TANH, r16.data_ptr(), o16.data_ptr(), 0, None, B, st), "tc")
    if which in ("all", "step"):
        diff._reverse_step(x_t, x0, eps, 5)
torch.cuda.synchronize()
print("done")
