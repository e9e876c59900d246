"""Stress: repeated perturbed-weight sampling loops (sample_mix / attention-mix kernels) with a result check between runs."""
import sys, torch
sys.path.insert(0, '.')
import skeletondiffusion_b200 as sdb
from skeletondiffusion_b200.testing import synth_state_dict
dev = torch.device('cuda:0')
spec = sdb.get_skeleton('amass')
ae, diff = sdb.build_models(spec, "cpu")
if "--identity" not in sys.argv:
    diff.load_state_dict(synth_state_dict(diff.state_dict(), seed=1, mode="perturbed", gain=2.5))
diff = diff.to(dev).eval()
diff.precision = "bf16x3"
W, S = 512, 50
z = torch.tanh(torch.randn(W, spec.num_nodes, 96, device=dev))
start = torch.randn(W * S, spec.num_nodes, 96, device=dev)
noise = torch.randn(W * S, 9, spec.num_nodes, 96, device=dev)
ref = None
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20
bad = 0
for i in range(n):
    lat, _ = diff.sample(batch_size=W * S, x_cond=z, start_noise=start, sampling_noise=noise)
    torch.cuda.synchronize()
    if ref is None:
        ref = lat.clone()
    else:
        d = (lat - ref).abs().max().item()
        if d != 0.0:
            bad += 1
            print(f"iteration {i}: result differs from iteration 0 by {d:.3e}")
print("stress done", n, "differing iterations:", bad)
