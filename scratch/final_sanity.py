"""Ten-second sanity of the built library on a B200: fp16x2 pipeline (fused GRU step, row norms from the epilogue, direct stores)
against the exact-fp32 path on the same inputs, identity and dense influence."""
import sys, torch
sys.path.insert(0, '.')
import skeletondiffusion_b200 as sdb
from skeletondiffusion_b200.testing import synth_state_dict
dev = torch.device('cuda:0')
spec = sdb.get_skeleton('amass')
for dense in (False, True):
    res = {}
    for prec in ('fp32', 'fp16x2'):
        ae, diff = sdb.build_models(spec, 'cpu', precision=prec)
        if dense:
            diff.load_state_dict(synth_state_dict(diff.state_dict(), seed=1, mode='perturbed', gain=1.0))
            ae.load_state_dict(synth_state_dict(ae.state_dict(), seed=2, mode='perturbed', gain=1.0))
        ae, diff = ae.to(dev).eval(), diff.to(dev).eval()
        g = torch.Generator().manual_seed(3)
        W, S, ph = 5, 31, 6
        obs = (torch.randn(W, spec.obs_length, spec.num_nodes, 3, generator=g) * 0.3).clamp(-1, 1).to(dev)
        start = torch.randn(W * S, spec.num_nodes, 96, generator=g).to(dev)
        noise = torch.randn(W * S, 9, spec.num_nodes, 96, generator=g).to(dev)
        res[prec] = sdb.get_prediction(obs, (ae, diff), num_samples=S, pred_length=ph, diffusion_conditioning=True,
                                       sampler_kwargs=dict(start_noise=start, sampling_noise=noise))
    err = float((res['fp16x2'] - res['fp32']).abs().max() / res['fp32'].abs().max())
    print('dense' if dense else 'identity', 'max rel diff fp16x2 vs fp32: %.2e' % err)
    assert err < 1e-4
print('ok')
