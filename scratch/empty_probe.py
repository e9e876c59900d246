"""What happens with a batch of zero rows (the reference returns empty tensors)?"""
import sys, torch, traceback
sys.path.insert(0, '.')
import skeletondiffusion_b200 as sdb
dev = torch.device('cuda:0')
spec = sdb.get_skeleton('h36m')
ae, diff = sdb.build_models(spec, dev)
N = spec.num_nodes
def attempt(name, fn):
    try:
        r = fn()
        print(f"{name}: ok {tuple(r.shape) if hasattr(r, 'shape') else r}")
    except Exception as e:
        print(f"{name}: {type(e).__name__}: {str(e)[:160]}")
x = torch.zeros(0, N, 96, device=dev)
cond = torch.zeros(0, N, 96, device=dev)
t = torch.zeros(0, dtype=torch.long, device=dev)
for prec in ("fp32", "bf16x3", "bf16"):
    attempt(f"denoiser[{prec}]", lambda: diff.model(x, t, None, cond, precision=prec))
    lay = diff.model.layers[0][0].block2.proj
    attempt(f"glin[{prec}]", lambda: lay.plan().forward(torch.zeros(0, N, 192, device=dev), precision=prec))
obs = torch.zeros(0, spec.obs_length, N, 3, device=dev)
attempt("encode", lambda: ae.get_past_embedding(obs))
attempt("decode", lambda: ae.decode(obs, x, None, ph=5))
attempt("sample", lambda: diff.sample(batch_size=0, x_cond=cond)[0])
attempt("get_prediction", lambda: sdb.get_prediction(obs, (ae, diff), num_samples=4, pred_length=5, diffusion_conditioning=True))
