import sys, torch
sys.path.insert(0, '.')
import skeletondiffusion_b200 as sdb
dev = torch.device('cuda:0')
spec = sdb.get_skeleton('amass')
ae, diff = sdb.build_models(spec, dev)
W, S, ph = 512, 50, 8
prec = sys.argv[1] if len(sys.argv) > 1 else None
obs = (torch.randn(W, spec.obs_length, spec.num_nodes, 3, device=dev) * 0.3).clamp(-1, 1)
lat = torch.tanh(torch.randn(W * S, spec.num_nodes, 96, device=dev))
for _ in range(2):
    out = ae.decode(obs, lat, None, ph=ph, precision=prec)
torch.cuda.synchronize()
print("ok", out.shape)
