"""One warm + one measured get_prediction at the AMASS eval batch (for ncu launch lists): python scratch/one_step.py [--perturbed]"""
import sys, torch
sys.path.insert(0, '.')
import skeletondiffusion_b200 as sdb
dev = torch.device('cuda:0')
spec = sdb.get_skeleton('amass')
ae, diff = sdb.build_models(spec, "cpu")
if "--perturbed" in sys.argv:
    from skeletondiffusion_b200.testing import synth_state_dict
    diff.load_state_dict(synth_state_dict(diff.state_dict(), seed=1, mode="perturbed", gain=2.5))
    ae.load_state_dict(synth_state_dict(ae.state_dict(), seed=2, mode="perturbed", gain=2.5))
ae, diff = ae.to(dev).eval(), diff.to(dev).eval()
import os
diff.precision = os.environ.get("SKELDIFF_PRECISION", "fp16x2")
W, S, ph = 512, 50, spec.pred_length
obs = (torch.randn(W, spec.obs_length, spec.num_nodes, 3, device=dev) * 0.3).clamp(-1, 1)
for _ in range(2):
    p = sdb.get_prediction(obs, (ae, diff), num_samples=S, pred_length=ph, diffusion_conditioning=True)
torch.cuda.synchronize()
print("ok", tuple(p.shape))
