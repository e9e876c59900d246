"""Times the three phases of get_prediction (encode / sample / decode) at the AMASS eval batch for each precision."""
import sys, torch
sys.path.insert(0, '.')
import skeletondiffusion_b200 as sdb
dev = torch.device('cuda:0')
spec = sdb.get_skeleton('amass')
ae, diff = sdb.build_models(spec, "cpu")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
if "--perturbed" in sys.argv:      # dense graph-influence matrices (trained-model-like), tests' stress weights
    from skeletondiffusion_b200.testing import synth_state_dict
    diff.load_state_dict(synth_state_dict(diff.state_dict(), seed=1, mode="perturbed", gain=2.5))
    ae.load_state_dict(synth_state_dict(ae.state_dict(), seed=2, mode="perturbed", gain=2.5))
ae, diff = ae.to(dev).eval(), diff.to(dev).eval()
W, S, ph = 512, 50, spec.pred_length
obs = (torch.randn(W, spec.obs_length, spec.num_nodes, 3, device=dev) * 0.3).clamp(-1, 1)

def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r

for prec in args or ["fp32", "bf16x3", "bf16"]:
    diff.precision = prec
    t_enc, z = timed(lambda: ae.get_past_embedding(obs, precision=prec if prec != 'bf16' else None))
    t_smp, (lat, _) = timed(lambda: diff.sample(batch_size=W * S, x_cond=z))
    t_dec, _ = timed(lambda: ae.decode(obs, lat, None, ph=ph, precision='bf16x3' if prec == 'bf16' else prec))
    print(f"{prec}: encode {t_enc:.1f} ms, sample {t_smp:.1f} ms, decode {t_dec:.1f} ms")
